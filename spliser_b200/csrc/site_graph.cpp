// Host-side construction of the site table and competing-site graph.
//
// Two regimes (SURVEY.md 8(a) "edge regimes"):
//   clean  -- unstranded run, or stranded run whose junction rows all carry '+' or '-': a site is
//             identified by position (resp. position + strand), so a hash lookup reproduces
//             binary_site_search (SpliSER_v0_1_8.py:175-225) exactly;
//   dirty  -- stranded run with other strand bytes (regtools '?'), or degenerate rows: the outcome
//             depends on where the reference's hand-rolled bisection lands in the current list, so
//             the list and the bisection are emulated step by step (S:186-221, S:347-350 with
//             Site.__lt__ of Gene_Site_Iter_Graph_v0_1_8.py:123-136).
// Only the *structure* is built here; alpha and PartnerCounts values are reduced on the GPU from
// the junction scores using the inc_* / einc_* segments.
#include "site_graph.h"

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <unordered_map>

namespace spl {
namespace {

inline bool is_pm(uint8_t s) { return s == '+' || s == '-'; }

struct PtKey {
    uint32_t a, pos;
    uint8_t strand;
    bool operator==(const PtKey& o) const { return a == o.a && pos == o.pos && strand == o.strand; }
};
struct PtKeyHash {
    size_t operator()(const PtKey& k) const {
        uint64_t h = ((uint64_t)k.a << 32 | k.pos) * 0x9E3779B97F4A7C15ull;
        return (size_t)(h ^ (h >> 29) ^ ((uint64_t)k.strand << 17));
    }
};

struct Builder {
    bool stranded = false;
    bool emulate = false;
    std::vector<int32_t> s_chrom, s_pos;
    std::vector<uint8_t> s_strand;
    std::vector<int64_t> s_line;
    std::vector<std::vector<int32_t>> order;            // emulate: per-chrom list of site ids
    std::unordered_map<uint64_t, int32_t> index;        // clean: key -> site id
    // edges in creation order
    std::vector<int32_t> pt_a, pt_b;                    // Partners entries (a has partner object b)
    std::vector<int32_t> pc_a, pc_pos;                  // PartnerCounts keys
    std::unordered_map<uint64_t, int32_t> pc_index;     // (a, pos) -> edge id
    std::unordered_map<PtKey, char, PtKeyHash> pt_seen;  // (a, b.pos, b.strand) per Site.__eq__
    std::vector<int32_t> inc_site, inc_line, einc_edge, einc_line;

    uint64_t key(int32_t c, int32_t p, uint8_t st) const {
        uint64_t sb = (stranded && st == '-') ? 1u : 0u;
        return ((uint64_t)(uint32_t)c << 34) | ((uint64_t)(uint32_t)p << 1) | sb;
    }

    // strand compatibility test shared by S:199 and S:204
    bool strand_ok(uint8_t q, uint8_t s) const { return q == s || !stranded || !is_pm(q); }

    // binary_site_search, S:175-225 (control flow kept: the result depends on it in the dirty regime)
    int32_t search(const std::vector<int32_t>& arr, int32_t pos, uint8_t strand) const {
        const int64_t length = (int64_t)arr.size();
        if (length == 0) return -1;
        int64_t idx = length / 2, past_max = length, past_min = 0, last_idx = -1, new_idx = idx;
        bool stuck = false, found = false;
        while (!stuck && !found) {
            const int32_t p = s_pos[arr[idx]];
            if (pos == p) {
                if (strand_ok(strand, s_strand[arr[idx]])) {
                    found = true;
                    break;
                }
                const int64_t cand[2] = {idx - 1, idx + 1};   // both computed from the landing index (S:203)
                for (int k = 0; k < 2; ++k) {
                    const int64_t a = cand[k];
                    if (a >= 0 && a < length && pos == s_pos[arr[a]] && strand_ok(strand, s_strand[arr[a]])) {
                        idx = a;
                        found = true;
                    }
                }
                break;
            } else if (pos >= p) {
                new_idx = idx + ((past_max - idx) / 2);
                past_min = idx;
            } else {
                new_idx = idx - ((idx - past_min) / 2);
                past_max = idx;
                if (idx == 1) new_idx = 0;
            }
            if (idx != last_idx) {
                last_idx = idx;
                idx = new_idx;
            } else {
                stuck = true;
            }
        }
        return found ? (int32_t)idx : -1;
    }

    // Site.__lt__, G:123-136 (a fall-through None is falsy)
    bool lt(int32_t a, int32_t b) const {
        if (stranded && s_pos[a] == s_pos[b]) {
            if (s_strand[a] == s_strand[b]) return false;
            return s_strand[a] == '+' && s_strand[b] == '-';
        }
        return s_pos[a] < s_pos[b];
    }

    void insort(std::vector<int32_t>& arr, int32_t x) const {   // bisect.insort_right, S:347-350
        size_t lo = 0, hi = arr.size();
        while (lo < hi) {
            size_t mid = (lo + hi) / 2;
            if (lt(x, arr[mid])) hi = mid; else lo = mid + 1;
        }
        arr.insert(arr.begin() + (ptrdiff_t)lo, x);
    }

    int32_t new_site(int32_t c, int32_t p, uint8_t st, int64_t line) {
        s_chrom.push_back(c);
        s_pos.push_back(p);
        s_strand.push_back(st);
        s_line.push_back(line);
        return (int32_t)s_pos.size() - 1;
    }

    void link(int32_t a, int32_t b, int32_t line) {
        // addPartner (G:260-262): `b not in a.Partners` is identity-or-__eq__; __eq__ (G:151-163) is
        // position equality when unstranded, position + strand-byte equality when stranded
        const PtKey pk{(uint32_t)a, (uint32_t)s_pos[b], stranded ? s_strand[b] : (uint8_t)0};
        if (pt_seen.emplace(pk, 1).second) {
            pt_a.push_back(a);
            pt_b.push_back(b);
        }
        // addPartnerCount (G:243-246): keyed by position only
        const uint64_t ck = ((uint64_t)(uint32_t)a << 32) | (uint32_t)s_pos[b];
        auto ce = pc_index.find(ck);
        int32_t e;
        if (ce == pc_index.end()) {
            e = (int32_t)pc_a.size();
            pc_index.emplace(ck, e);
            pc_a.push_back(a);
            pc_pos.push_back(s_pos[b]);
        } else {
            e = ce->second;
        }
        einc_edge.push_back(e);
        einc_line.push_back(line);
    }
};

// stable counting sort of `n` items with integer keys in [0, n_keys): returns offsets and fills perm
void group_by(const std::vector<int32_t>& keys, int64_t n_keys, std::vector<int64_t>& off, std::vector<int32_t>& perm) {
    off.assign((size_t)n_keys + 1, 0);
    for (int32_t k : keys) off[(size_t)k + 1]++;
    for (int64_t i = 0; i < n_keys; ++i) off[(size_t)i + 1] += off[(size_t)i];
    perm.resize(keys.size());
    std::vector<int64_t> cur(off.begin(), off.end() - 1);
    for (size_t i = 0; i < keys.size(); ++i) perm[(size_t)cur[(size_t)keys[i]]++] = (int32_t)i;
}

void build_rp(SiteGraph& g) {
    // anchor = first site index of the chromosome whose position equals the partner position
    std::vector<int32_t> anchor_of;   // per (t, e) pair
    std::vector<int32_t> t_of;
    anchor_of.reserve(g.pc_pos.size());
    t_of.reserve(g.pc_pos.size());
    const bool direct = !g.dirty_regime && g.pt_site.size() == g.pc_pos.size() && g.gap_index.empty();
    std::vector<int32_t> first_at;    // first site index sharing (chrom, pos) with site s
    if (direct) {
        first_at.resize((size_t)g.n_sites);
        for (int64_t s = 0; s < g.n_sites; ++s)
            first_at[(size_t)s] = (s > 0 && g.chrom[(size_t)s - 1] == g.chrom[(size_t)s] && g.pos[(size_t)s - 1] == g.pos[(size_t)s])
                                      ? first_at[(size_t)s - 1] : (int32_t)s;
    }
    for (int64_t t = 0; t < g.n_sites; ++t) {
        const int32_t c = g.chrom[(size_t)t];
        const int32_t* lo = g.pos.data() + g.cs_off[(size_t)c];
        const int32_t* hi = g.pos.data() + g.cs_off[(size_t)c + 1];
        for (int64_t e = g.pc_off[(size_t)t]; e < g.pc_off[(size_t)t + 1]; ++e) {
            if (direct) {                                              // PartnerCounts entry e <-> partner object pt_site[e]
                anchor_of.push_back(first_at[(size_t)g.pt_site[(size_t)e]]);
                t_of.push_back((int32_t)t);
                continue;
            }
            const int32_t* it = std::lower_bound(lo, hi, g.pc_pos[(size_t)e]);
            if (it == hi || *it != g.pc_pos[(size_t)e]) continue;   // cannot happen: partners are rows
            anchor_of.push_back((int32_t)(it - g.pos.data()));
            t_of.push_back((int32_t)t);
        }
    }
    std::vector<int32_t> perm;
    group_by(anchor_of, g.n_sites, g.rp_off, perm);
    g.rp_site.resize(perm.size());
    for (size_t i = 0; i < perm.size(); ++i) g.rp_site[i] = t_of[(size_t)perm[i]];
}

}  // namespace

namespace {
void build_rp(SiteGraph& g);

// competitors: positions of partners-of-partners other than own position (S:364-372), then the reverse index
void finish_graph(SiteGraph& g) {
    const int64_t S = g.n_sites;
    g.cp_off.assign((size_t)S + 1, 0);
    g.cp_pos.clear();
    std::vector<int32_t> tmp;
    for (int64_t t = 0; t < S; ++t) {
        tmp.clear();
        for (int64_t a = g.pt_off[(size_t)t]; a < g.pt_off[(size_t)t + 1]; ++a) {
            const int32_t p = g.pt_site[(size_t)a];
            for (int64_t q = g.pt_off[(size_t)p]; q < g.pt_off[(size_t)p + 1]; ++q) {
                const int32_t cpos = g.pos[(size_t)g.pt_site[(size_t)q]];
                if (cpos != g.pos[(size_t)t]) tmp.push_back(cpos);
            }
        }
        if (tmp.size() > 1) {
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        }
        g.cp_pos.insert(g.cp_pos.end(), tmp.begin(), tmp.end());
        g.cp_off[(size_t)t + 1] = (int64_t)g.cp_pos.size();
    }
    build_rp(g);
}
}  // namespace

namespace {

// stable LSD radix sort of (key, val) pairs, 8-bit digits; digits on which all keys agree are skipped
void radix_sort_pairs(std::vector<uint64_t>& key, std::vector<uint32_t>& val, int key_bits) {
    const size_t n = key.size();
    if (n < 2) return;
    uint64_t all_or = 0, all_and = ~0ull;
    bool sorted = true;
    for (size_t i = 0; i < n; ++i) {
        all_or |= key[i]; all_and &= key[i];
        if (i && key[i] < key[i - 1]) sorted = false;
    }
    if (sorted) return;
    const uint64_t differ = all_or ^ all_and;
    std::vector<uint64_t> k2(n);
    std::vector<uint32_t> v2(n);
    size_t cnt[256];
    for (int shift = 0; shift < key_bits; shift += 8) {
        if (!((differ >> shift) & 0xff)) continue;
        std::fill(cnt, cnt + 256, (size_t)0);
        for (size_t i = 0; i < n; ++i) cnt[(key[i] >> shift) & 0xff]++;
        size_t run = 0;
        for (size_t b = 0; b < 256; ++b) { const size_t c = cnt[b]; cnt[b] = run; run += c; }
        const uint64_t* ks = key.data(); const uint32_t* vs = val.data();
        uint64_t* kd = k2.data(); uint32_t* vd = v2.data();
        for (size_t i = 0; i < n; ++i) {
            const size_t d = cnt[(ks[i] >> shift) & 0xff]++;
            kd[d] = ks[i]; vd[d] = vs[i];
        }
        key.swap(k2); val.swap(v2);
    }
}

// LSD radix sort of packed 64-bit words on bits [lo_bit, hi_bit); lower bits ride along (stable)
void radix_sort_packed(std::vector<uint64_t>& a, int lo_bit, int hi_bit) {
    const size_t n = a.size();
    if (n < 2) return;
    uint64_t all_or = 0, all_and = ~0ull;
    bool sorted = true;
    const uint64_t hmask = hi_bit >= 64 ? ~0ull : (((uint64_t)1 << hi_bit) - 1);
    for (size_t i = 0; i < n; ++i) {
        all_or |= a[i]; all_and &= a[i];
        if (i && ((a[i] & hmask) >> lo_bit) < ((a[i - 1] & hmask) >> lo_bit)) sorted = false;
    }
    if (sorted) return;
    const uint64_t differ = all_or ^ all_and;
    std::vector<uint64_t> b(n);
    size_t cnt[256];
    for (int shift = lo_bit; shift < hi_bit; shift += 8) {
        const int width = std::min(8, hi_bit - shift);
        const uint64_t dm = ((uint64_t)1 << width) - 1;
        if (!((differ >> shift) & dm)) continue;
        std::fill(cnt, cnt + 256, (size_t)0);
        for (size_t i = 0; i < n; ++i) cnt[(a[i] >> shift) & dm]++;
        size_t run = 0;
        for (size_t k = 0; k < 256; ++k) { const size_t c = cnt[k]; cnt[k] = run; run += c; }
        const uint64_t* src = a.data(); uint64_t* dst = b.data();
        for (size_t i = 0; i < n; ++i) dst[cnt[(src[i] >> shift) & dm]++] = src[i];
        a.swap(b);
    }
}

int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

// Clean regime: a site is (chrom, pos[, '+'/'-']); everything is a sort / unique / group-by over the junction rows.
// Produces exactly what the line-by-line construction of the reference produces (list order, first-seen
// strand and line, Partners / PartnerCounts in first-appearance order).
void build_clean_sorted(int32_t n_chrom, int64_t n_junc, const int32_t* j_chrom, const int32_t* j_left, const int32_t* j_right,
                        const uint8_t* j_strand, bool stranded, SiteGraph& g) {
    const size_t J = (size_t)n_junc;
    g = SiteGraph();
    g.n_chrom = n_chrom;
    const char* tenv = std::getenv("SPLISER_TIMING");
    const bool timing = tenv && tenv[0] == '2';
    auto tprev = std::chrono::steady_clock::now();
#define TICK(name) do { if (timing) { auto tn = std::chrono::steady_clock::now(); fprintf(stderr, "  %-16s %.2f ms\n", name, std::chrono::duration<double, std::milli>(tn - tprev).count()); tprev = tn; } } while (0)
    // ---- endpoints -> sites
    std::vector<uint64_t> key(2 * J);
    std::vector<uint32_t> val(2 * J);
    uint64_t kmax = 0;
    for (size_t i = 0; i < J; ++i) {
        const uint64_t sb = (stranded && j_strand[i] == '-') ? 1u : 0u;
        const uint64_t c = (uint64_t)(uint32_t)j_chrom[i] << 33;
        key[2 * i] = c | ((uint64_t)(uint32_t)j_left[i] << 1) | sb;
        key[2 * i + 1] = c | ((uint64_t)(uint32_t)j_right[i] << 1) | sb;
        val[2 * i] = (uint32_t)(2 * i); val[2 * i + 1] = (uint32_t)(2 * i + 1);
        kmax |= key[2 * i] | key[2 * i + 1];
    }
    TICK("endpoint keys");
    const int vb = bits_for(2 * J), kb = bits_for(kmax);
    if (kb + vb <= 64) {                 // usual case: key and row id share one word -> one array to move
        for (size_t e = 0; e < 2 * J; ++e) key[e] = (key[e] << vb) | val[e];
        radix_sort_packed(key, vb, vb + kb);
        const uint64_t vm = ((uint64_t)1 << vb) - 1;
        for (size_t e = 0; e < 2 * J; ++e) { val[e] = (uint32_t)(key[e] & vm); key[e] >>= vb; }
    } else {
        radix_sort_pairs(key, val, kb);
    }
    TICK("endpoint sort");
    std::vector<int32_t> site_of(2 * J);
    g.chrom.reserve(2 * J); g.pos.reserve(2 * J); g.strand.reserve(2 * J); g.first_line.reserve(2 * J);
    g.cls.reserve(2 * J); g.inc_off.reserve(2 * J + 1);
    g.cs_off.assign((size_t)n_chrom + 1, 0);
    g.inc_off.clear(); g.inc_line.resize(2 * J);
    int64_t S = 0;
    for (size_t e = 0; e < 2 * J; ++e) {
        if (e == 0 || key[e] != key[e - 1]) {
            const size_t line = val[e] >> 1;
            g.chrom.push_back(j_chrom[line]);
            g.pos.push_back((int32_t)((key[e] >> 1) & 0xffffffffu));
            g.strand.push_back(j_strand[line]);                       // first-seen strand (stable sort keeps row order)
            g.first_line.push_back((int64_t)line);
            g.cls.push_back(!stranded ? CLS_ANY : j_strand[line] == '+' ? CLS_PLUS : j_strand[line] == '-' ? CLS_MINUS : CLS_NEVER);
            g.inc_off.push_back((int64_t)e);
            g.cs_off[(size_t)j_chrom[line] + 1]++;
            ++S;
        }
        site_of[val[e]] = (int32_t)(S - 1);
        g.inc_line[e] = (int32_t)(val[e] >> 1);
    }
    g.inc_off.push_back((int64_t)(2 * J));
    g.n_sites = S;
    for (int32_t c = 0; c < n_chrom; ++c) g.cs_off[(size_t)c + 1] += g.cs_off[(size_t)c];
    TICK("sites");
    // ---- directed edges -> PartnerCounts entries
    const int sbits = bits_for((uint64_t)(S > 0 ? S - 1 : 0));
    for (size_t i = 0; i < J; ++i) {
        const uint64_t a = (uint64_t)(uint32_t)site_of[2 * i], b = (uint64_t)(uint32_t)site_of[2 * i + 1];
        key[2 * i] = (a << 32) | b; key[2 * i + 1] = (b << 32) | a;
        val[2 * i] = (uint32_t)i; val[2 * i + 1] = (uint32_t)i;       // row order == first-appearance order
    }
    TICK("edge keys");
    const int lb = bits_for(J);
    if (32 + sbits + lb <= 64) {
        for (size_t e = 0; e < 2 * J; ++e) key[e] = (key[e] << lb) | val[e];
        radix_sort_packed(key, lb, lb + 32 + sbits);
        const uint64_t lm = ((uint64_t)1 << lb) - 1;
        for (size_t e = 0; e < 2 * J; ++e) { val[e] = (uint32_t)(key[e] & lm); key[e] >>= lb; }
    } else {
        radix_sort_pairs(key, val, 32 + sbits);
    }
    TICK("edge sort");
    std::vector<uint64_t> ukey;       // (src << 32) | first_line   -> order of Partners inside a site
    std::vector<uint32_t> uidx;       // index of the unique edge
    std::vector<int32_t> u_dst;
    std::vector<int64_t> u_lo;        // segment [u_lo[u], u_lo[u+1]) of the sorted rows
    ukey.reserve(2 * J); uidx.reserve(2 * J); u_dst.reserve(2 * J); u_lo.reserve(2 * J + 1);
    for (size_t e = 0; e < 2 * J; ++e) {
        if (e == 0 || key[e] != key[e - 1]) {
            ukey.push_back((key[e] & 0xffffffff00000000ull) | val[e]);
            uidx.push_back((uint32_t)u_dst.size());
            u_dst.push_back((int32_t)(key[e] & 0xffffffffu));
            u_lo.push_back((int64_t)e);
        }
    }
    u_lo.push_back((int64_t)(2 * J));
    const size_t E = u_dst.size();
    TICK("unique edges");
    const int eb = bits_for(E);
    if (32 + sbits + eb <= 64) {
        for (size_t x = 0; x < E; ++x) ukey[x] = (ukey[x] << eb) | uidx[x];
        radix_sort_packed(ukey, eb, eb + 32 + sbits);
        const uint64_t em = ((uint64_t)1 << eb) - 1;
        for (size_t x = 0; x < E; ++x) { uidx[x] = (uint32_t)(ukey[x] & em); ukey[x] >>= eb; }
    } else {
        radix_sort_pairs(ukey, uidx, 32 + sbits);
    }
    TICK("edge order sort");
    g.pt_off.assign((size_t)S + 1, 0);
    g.pt_site.resize(E); g.pc_pos.resize(E);
    g.einc_off.assign(E + 1, 0);
    g.einc_line.resize(2 * J);
    int64_t w = 0;
    for (size_t x = 0; x < E; ++x) {
        const uint32_t u = uidx[x];
        const size_t src = (size_t)(ukey[x] >> 32);
        g.pt_off[src + 1]++;
        g.pt_site[x] = u_dst[u];
        g.pc_pos[x] = g.pos[(size_t)u_dst[u]];
        for (int64_t e = u_lo[u]; e < u_lo[u + 1]; ++e) g.einc_line[(size_t)w++] = (int32_t)val[(size_t)e];
        g.einc_off[x + 1] = w;
    }
    for (int64_t t = 0; t < S; ++t) g.pt_off[(size_t)t + 1] += g.pt_off[(size_t)t];
    g.pc_off = g.pt_off;              // clean regime: one PartnerCounts key per partner object
    TICK("csr");
}


// Sites of different chromosomes never interact (the chromosome is the top sort key and partners share the
// junction's chromosome), so the clean build runs one independent sub-problem per chromosome on a small pool of
// host threads and concatenates the pieces with shifted indices.
void build_clean_by_chromosome(int32_t n_chrom, int64_t n_junc, const int32_t* j_chrom, const int32_t* j_left, const int32_t* j_right,
                               const uint8_t* j_strand, bool stranded, SiteGraph& g) {
    const size_t J = (size_t)n_junc, NC = (size_t)n_chrom;
    std::vector<int64_t> roff(NC + 1, 0);
    for (size_t i = 0; i < J; ++i) roff[(size_t)j_chrom[i] + 1]++;
    size_t n_used = 0;
    for (size_t c = 0; c < NC; ++c) { n_used += roff[c + 1] != 0; roff[c + 1] += roff[c]; }
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* e = std::getenv("SPLISER_GRAPH_THREADS")) hw = (unsigned)std::max(1, atoi(e));
    if (n_used <= 1 || J < 4096 || hw <= 1) {
        build_clean_sorted(n_chrom, n_junc, j_chrom, j_left, j_right, j_strand, stranded, g);
        finish_graph(g);
        return;
    }
    // rows of each chromosome, in table order (keeps first-appearance semantics)
    std::vector<int32_t> rows(J), lft(J), rgt(J);
    std::vector<uint8_t> strd(J);
    {
        std::vector<int64_t> cur(roff.begin(), roff.end() - 1);
        for (size_t i = 0; i < J; ++i) {
            const size_t d = (size_t)cur[(size_t)j_chrom[i]]++;
            rows[d] = (int32_t)i; lft[d] = j_left[i]; rgt[d] = j_right[i]; strd[d] = j_strand[i];
        }
    }
    std::vector<int32_t> order;                      // chromosomes with rows, biggest first
    for (size_t c = 0; c < NC; ++c) if (roff[c + 1] > roff[c]) order.push_back((int32_t)c);
    std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return roff[(size_t)a + 1] - roff[(size_t)a] > roff[(size_t)b + 1] - roff[(size_t)b];
    });
    const bool timing = std::getenv("SPLISER_TIMING") != nullptr;
    auto tp = std::chrono::steady_clock::now();
    auto lap = [&](const char* name) {
        if (!timing) return;
        const auto tn = std::chrono::steady_clock::now();
        fprintf(stderr, "  [by-chrom] %-10s %.2f ms\n", name, std::chrono::duration<double, std::milli>(tn - tp).count());
        tp = tn;
    };
    std::vector<SiteGraph> piece(NC);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= order.size()) return;
            const size_t c = (size_t)order[k];
            const size_t r0 = (size_t)roff[c], n = (size_t)(roff[c + 1] - roff[c]);
            std::vector<int32_t> zero(n, 0);
            const auto w0 = std::chrono::steady_clock::now();
            build_clean_sorted(1, (int64_t)n, zero.data(), lft.data() + r0, rgt.data() + r0, strd.data() + r0, stranded, piece[c]);
            const auto w1 = std::chrono::steady_clock::now();
            finish_graph(piece[c]);
            if (timing) {
                const auto w2 = std::chrono::steady_clock::now();
                fprintf(stderr, "    chrom %zu rows %zu: start +%.2f sorted %.2f finish %.2f ms\n", c, n, std::chrono::duration<double, std::milli>(w0 - tp).count(),
                        std::chrono::duration<double, std::milli>(w1 - w0).count(), std::chrono::duration<double, std::milli>(w2 - w1).count());
            }
        }
    };
    const size_t nt = std::min<size_t>(std::min<size_t>(hw, 16), order.size());
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nt; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    lap("pieces");
    // ---- concatenate
    g = SiteGraph();
    g.n_chrom = n_chrom;
    g.cs_off.assign(NC + 1, 0);
    size_t S = 0, E = 0, C = 0, RP = 0;
    for (size_t c = 0; c < NC; ++c) {
        const SiteGraph& p = piece[c];
        g.cs_off[c] = (int64_t)S;
        S += (size_t)p.n_sites; E += p.pc_pos.size(); C += p.cp_pos.size(); RP += p.rp_site.size();
    }
    g.cs_off[NC] = (int64_t)S;
    g.n_sites = (int64_t)S;
    g.chrom.resize(S); g.pos.resize(S); g.strand.resize(S); g.cls.resize(S); g.first_line.resize(S);
    g.pt_off.resize(S + 1); g.pt_site.resize(E); g.pc_pos.resize(E);
    g.cp_off.resize(S + 1); g.cp_pos.resize(C);
    g.rp_off.resize(S + 1); g.rp_site.resize(RP);
    g.inc_off.resize(S + 1); g.inc_line.resize(2 * J);
    g.einc_off.resize(E + 1); g.einc_line.resize(2 * J);
    std::vector<size_t> s0(NC + 1, 0), e0(NC + 1, 0), c0(NC + 1, 0), q0(NC + 1, 0), l0(NC + 1, 0);
    for (size_t c = 0; c < NC; ++c) {
        const SiteGraph& p = piece[c];
        s0[c + 1] = s0[c] + (size_t)p.n_sites; e0[c + 1] = e0[c] + p.pc_pos.size(); c0[c + 1] = c0[c] + p.cp_pos.size();
        q0[c + 1] = q0[c] + p.rp_site.size(); l0[c + 1] = l0[c] + p.inc_line.size();
    }
    std::atomic<size_t> nextm{0};
    auto merger = [&]() {
        for (;;) {
            const size_t k = nextm.fetch_add(1);
            if (k >= order.size()) return;
            const size_t c = (size_t)order[k];
            const SiteGraph& p = piece[c];
            const size_t n = (size_t)p.n_sites, so = s0[c], eo = e0[c], co = c0[c], qo = q0[c], lo = l0[c];
            const int32_t* rw = rows.data() + roff[c];
            for (size_t i = 0; i < n; ++i) {
                g.chrom[so + i] = (int32_t)c; g.pos[so + i] = p.pos[i]; g.strand[so + i] = p.strand[i]; g.cls[so + i] = p.cls[i];
                g.first_line[so + i] = rw[p.first_line[i]];
                g.pt_off[so + i] = p.pt_off[i] + (int64_t)eo; g.cp_off[so + i] = p.cp_off[i] + (int64_t)co;
                g.rp_off[so + i] = p.rp_off[i] + (int64_t)qo; g.inc_off[so + i] = p.inc_off[i] + (int64_t)lo;
            }
            for (size_t x = 0; x < p.pc_pos.size(); ++x) {
                g.pt_site[eo + x] = p.pt_site[x] + (int32_t)so; g.pc_pos[eo + x] = p.pc_pos[x];
                g.einc_off[eo + x] = p.einc_off[x] + (int64_t)lo;
            }
            for (size_t x = 0; x < p.cp_pos.size(); ++x) g.cp_pos[co + x] = p.cp_pos[x];
            for (size_t x = 0; x < p.rp_site.size(); ++x) g.rp_site[qo + x] = p.rp_site[x] + (int32_t)so;
            for (size_t x = 0; x < p.inc_line.size(); ++x) g.inc_line[lo + x] = rw[p.inc_line[x]];
            for (size_t x = 0; x < p.einc_line.size(); ++x) g.einc_line[lo + x] = rw[p.einc_line[x]];
        }
    };
    pool.clear();
    for (size_t t = 1; t < nt; ++t) pool.emplace_back(merger);
    merger();
    for (auto& t : pool) t.join();
    lap("merge");
    g.pt_off[S] = (int64_t)E; g.cp_off[S] = (int64_t)C; g.rp_off[S] = (int64_t)RP; g.inc_off[S] = (int64_t)(2 * J);
    g.einc_off[E] = (int64_t)(2 * J);
    g.pc_off = g.pt_off;
}

}  // namespace

std::string build_site_graph(int32_t n_chrom, int64_t n_junc, const int32_t* j_chrom, const int32_t* j_left,
                             const int32_t* j_right, const uint8_t* j_strand, bool stranded, SiteGraph& g) {
    if (n_chrom < 0 || n_junc < 0) return "negative size";
    if (n_junc > 0 && (!j_chrom || !j_left || !j_right || !j_strand)) return "null junction array";
    if (n_junc >= (int64_t)1 << 30) return "too many junction rows";
    Builder b;
    b.stranded = stranded;
    for (int64_t i = 0; i < n_junc; ++i) {
        if (j_chrom[i] < 0 || j_chrom[i] >= n_chrom) return "junction chromosome index out of range";
        if ((stranded && !is_pm(j_strand[i])) || j_left[i] == j_right[i] || j_left[i] < 0 || j_right[i] < 0) b.emulate = true;
    }
    if (const char* f = std::getenv("SPLISER_FORCE_EMULATION")) {
        if (f[0] == '1') b.emulate = true;
    }
    if (!b.emulate) {
        const auto t0 = std::chrono::steady_clock::now();
        build_clean_by_chromosome(n_chrom, n_junc, j_chrom, j_left, j_right, j_strand, stranded, g);
        if (std::getenv("SPLISER_TIMING")) {
            const auto t2 = std::chrono::steady_clock::now();
            fprintf(stderr, "site graph: clean build %.2f ms\n", std::chrono::duration<double, std::milli>(t2 - t0).count());
        }
        return "";
    }
    b.order.resize((size_t)n_chrom);
    b.pc_index.reserve((size_t)n_junc * 2);
    b.pt_seen.reserve((size_t)n_junc * 2);

    for (int64_t i = 0; i < n_junc; ++i) {
        const int32_t c = j_chrom[i], l = j_left[i], r = j_right[i];
        const uint8_t st = j_strand[i];
        auto& arr = b.order[(size_t)c];
        const int32_t li = b.search(arr, l, st), ri = b.search(arr, r, st);   // both before any insert, S:291-292
        int32_t sl = li >= 0 ? arr[(size_t)li] : -1;
        int32_t sr = ri >= 0 ? arr[(size_t)ri] : -1;
        bool lnew = false, rnew = false;
        if (sl < 0) { sl = b.new_site(c, l, st, i); lnew = true; }
        if (sr < 0) { sr = b.new_site(c, r, st, i); rnew = true; }
        b.inc_site.push_back(sl); b.inc_line.push_back((int32_t)i);     // addAlphaCount, S:341
        b.inc_site.push_back(sr); b.inc_line.push_back((int32_t)i);
        if (lnew) b.insort(arr, sl);
        if (rnew) b.insort(arr, sr);
        b.link(sl, sr, (int32_t)i);                                       // S:352-355
        b.link(sr, sl, (int32_t)i);
    }

    // ---- creation id -> list order --------------------------------------------------------------
    const int64_t S = (int64_t)b.s_pos.size();
    std::vector<int32_t> new_of((size_t)S), old_of((size_t)S);
    g = SiteGraph();
    g.n_chrom = n_chrom;
    g.n_sites = S;
    g.dirty_regime = true;
    g.cs_off.assign((size_t)n_chrom + 1, 0);
    {
        int64_t k = 0;
        for (int32_t c = 0; c < n_chrom; ++c) {
            g.cs_off[(size_t)c] = k;
            for (int32_t id : b.order[(size_t)c]) old_of[(size_t)k++] = id;
        }
        g.cs_off[(size_t)n_chrom] = k;
    }
    for (int64_t k = 0; k < S; ++k) new_of[(size_t)old_of[(size_t)k]] = (int32_t)k;
    g.chrom.resize((size_t)S); g.pos.resize((size_t)S); g.strand.resize((size_t)S);
    g.cls.resize((size_t)S); g.first_line.resize((size_t)S);
    for (int64_t k = 0; k < S; ++k) {
        const int32_t o = old_of[(size_t)k];
        g.chrom[(size_t)k] = b.s_chrom[o];
        g.pos[(size_t)k] = b.s_pos[o];
        g.strand[(size_t)k] = b.s_strand[o];
        g.first_line[(size_t)k] = b.s_line[o];
        g.cls[(size_t)k] = !stranded ? CLS_ANY : b.s_strand[o] == '+' ? CLS_PLUS : b.s_strand[o] == '-' ? CLS_MINUS : CLS_NEVER;
    }

    std::vector<int32_t> keys, perm;
    // Partners CSR
    keys.resize(b.pt_a.size());
    for (size_t i = 0; i < keys.size(); ++i) keys[i] = new_of[(size_t)b.pt_a[i]];
    group_by(keys, S, g.pt_off, perm);
    g.pt_site.resize(perm.size());
    for (size_t i = 0; i < perm.size(); ++i) g.pt_site[i] = new_of[(size_t)b.pt_b[(size_t)perm[i]]];
    // PartnerCounts CSR (+ old edge id -> new edge id)
    keys.resize(b.pc_a.size());
    for (size_t i = 0; i < keys.size(); ++i) keys[i] = new_of[(size_t)b.pc_a[i]];
    group_by(keys, S, g.pc_off, perm);
    const int64_t E = (int64_t)perm.size();
    std::vector<int32_t> new_edge((size_t)E);
    g.pc_pos.resize((size_t)E);
    for (int64_t i = 0; i < E; ++i) {
        new_edge[(size_t)perm[(size_t)i]] = (int32_t)i;
        g.pc_pos[(size_t)i] = b.pc_pos[(size_t)perm[(size_t)i]];
    }
    // reduction segments
    keys.resize(b.inc_site.size());
    for (size_t i = 0; i < keys.size(); ++i) keys[i] = new_of[(size_t)b.inc_site[i]];
    group_by(keys, S, g.inc_off, perm);
    g.inc_line.resize(perm.size());
    for (size_t i = 0; i < perm.size(); ++i) g.inc_line[i] = b.inc_line[(size_t)perm[i]];
    keys.resize(b.einc_edge.size());
    for (size_t i = 0; i < keys.size(); ++i) keys[i] = new_edge[(size_t)b.einc_edge[i]];
    group_by(keys, E, g.einc_off, perm);
    g.einc_line.resize(perm.size());
    for (size_t i = 0; i < perm.size(); ++i) g.einc_line[i] = b.einc_line[(size_t)perm[i]];
    finish_graph(g);
    return "";
}

std::string build_recount_graph(int32_t n_chrom, int64_t n_sites, const int32_t* s_chrom, const int32_t* s_pos,
                                const uint8_t* s_strand, const int64_t* p_off, const int32_t* p_pos,
                                const int64_t* c_off, const int32_t* c_pos, bool stranded, SiteGraph& g) {
    if (n_chrom < 0 || n_sites < 0) return "negative size";
    if (n_sites > 0 && (!s_chrom || !s_pos || !s_strand || !p_off || !c_off)) return "null site array";
    if (n_sites >= (int64_t)1 << 30) return "too many sites";
    struct Row { int32_t chrom, pos; int64_t gap; };
    std::vector<Row> rows;
    rows.reserve((size_t)n_sites * 2);
    std::unordered_map<uint64_t, char> have;
    have.reserve((size_t)n_sites * 2);
    auto pk = [](int32_t c, int32_t p) { return ((uint64_t)(uint32_t)c << 32) | (uint32_t)p; };
    for (int64_t i = 0; i < n_sites; ++i) {
        if (p_off[i + 1] < p_off[i] || c_off[i + 1] < c_off[i]) return "offsets not monotone";
        // a region unknown to this sample's BAM: the reference's samtools call yields no reads, so
        // the gap keeps the zeros the caller initialised its outputs with
        if (s_chrom[i] < 0 || s_chrom[i] >= n_chrom) continue;
        rows.push_back({s_chrom[i], s_pos[i], i});
        have.emplace(pk(s_chrom[i], s_pos[i]), 1);
    }
    const size_t n_real = rows.size();
    for (size_t k = 0; k < n_real; ++k) {
        const int64_t i = rows[k].gap;
        for (int64_t e = p_off[i]; e < p_off[i + 1]; ++e) {
            const uint64_t key = pk(rows[k].chrom, p_pos[e]);
            if (have.find(key) == have.end()) {
                have.emplace(key, 1);
                rows.push_back({rows[k].chrom, p_pos[e], -1});
            }
        }
    }
    std::stable_sort(rows.begin(), rows.end(), [](const Row& a, const Row& b) {
        if (a.chrom != b.chrom) return a.chrom < b.chrom;
        return a.pos < b.pos;
    });
    const int64_t S = (int64_t)rows.size();
    g = SiteGraph();
    g.n_chrom = n_chrom;
    g.n_sites = S;
    g.cs_off.assign((size_t)n_chrom + 1, 0);
    g.chrom.resize((size_t)S); g.pos.resize((size_t)S); g.strand.resize((size_t)S); g.cls.resize((size_t)S);
    g.first_line.assign((size_t)S, -1);
    g.gap_index.resize((size_t)S);
    g.pt_off.assign((size_t)S + 1, 0);
    g.inc_off.assign((size_t)S + 1, 0);
    g.pc_off.assign((size_t)S + 1, 0);
    g.cp_off.assign((size_t)S + 1, 0);
    std::vector<int32_t> tmp;
    for (int64_t k = 0; k < S; ++k) {
        const Row& r = rows[(size_t)k];
        g.cs_off[(size_t)r.chrom + 1]++;
        g.chrom[(size_t)k] = r.chrom;
        g.pos[(size_t)k] = r.pos;
        g.gap_index[(size_t)k] = r.gap;
        if (r.gap < 0) {
            g.strand[(size_t)k] = 0;
            g.cls[(size_t)k] = CLS_PSEUDO;
        } else {
            const uint8_t st = s_strand[r.gap];
            g.strand[(size_t)k] = st;
            g.cls[(size_t)k] = !stranded ? CLS_ANY : st == '+' ? CLS_PLUS : st == '-' ? CLS_MINUS : CLS_NEVER;
            tmp.assign(p_pos + p_off[r.gap], p_pos + p_off[r.gap + 1]);
            // keys of a dict: unique, order irrelevant for membership tests
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            g.pc_pos.insert(g.pc_pos.end(), tmp.begin(), tmp.end());
            tmp.assign(c_pos + c_off[r.gap], c_pos + c_off[r.gap + 1]);
            std::sort(tmp.begin(), tmp.end());
            tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            g.cp_pos.insert(g.cp_pos.end(), tmp.begin(), tmp.end());
        }
        g.pc_off[(size_t)k + 1] = (int64_t)g.pc_pos.size();
        g.cp_off[(size_t)k + 1] = (int64_t)g.cp_pos.size();
    }
    for (int32_t c = 0; c < n_chrom; ++c) g.cs_off[(size_t)c + 1] += g.cs_off[(size_t)c];
    g.einc_off.assign(g.pc_pos.size() + 1, 0);
    build_rp(g);
    return "";
}

}  // namespace spl
