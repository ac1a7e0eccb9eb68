// K1 on the device: site table + competing-site graph from the junction table (clean regime).
//
// Replaces the structural half of findAlphaCounts (SpliSER_v0_1_8.py:289-355: binary_site_search, insort, addPartner,
// addPartnerCount) and findCompetitorPos (S:364-372) by sort / unique / group-by passes over the junction rows:
//   A  endpoint keys (chrom, pos, strand bit | row id) -> radix sort -> unique = sites in the reference's list order
//      (chrom_index order, position ascending, '+' before '-' per Site.__lt__, G:123-136); the first row of a run
//      gives the displayed strand and the Gene lookup line (first-seen semantics, S:296-345)
//   B  directed (site, partner | row) keys -> radix sort -> unique = PartnerCounts entries with their row segments
//   C  (site, first row | entry) -> radix sort = Partners / PartnerCounts in first-appearance order (G:243-262)
//   D  (site, partner-of-partner position) candidates -> radix sort -> unique = CompetitorPos (G:264-266)
//   E  reverse-partner index + hot flags, direct-address bin index of the site table
// The radix sort is a plain stable LSD sort (8-bit digits; per-tile histogram, scan, ranked scatter) written here;
// tables are a few 10^5..10^6 keys, so launch count matters more than bandwidth.
// The dirty regime ('?' strands in a stranded run, l == r rows) keeps the sequential host emulation (site_graph.cpp).
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <string>

#include "device_types.h"
#include "graph_build.h"

namespace spl {

namespace {

constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint64_t* __restrict__ in, uint32_t n, int shift, uint32_t mask, uint32_t* __restrict__ hist, uint32_t n_tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(in[i] >> shift) & mask], 1u);
    }
    __syncthreads();
    hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];      // digit-major: one scan gives every (digit, tile) offset
}

// stable ranked scatter: element order inside a tile is (round, warp, lane) = index order
__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t* __restrict__ in, uint64_t* __restrict__ out, uint32_t n, int shift, uint32_t mask,
             const uint32_t* __restrict__ offs, uint32_t n_tiles) {
    __shared__ uint32_t wc[RS_THREADS / 32][256];
    __shared__ uint32_t base[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = offs[threadIdx.x * n_tiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) wc[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t t0 = blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ITEMS; ++r) {
        const uint32_t i = t0 + r * RS_THREADS + threadIdx.x;
        const bool live = i < n;
        const uint64_t key = live ? in[i] : 0ull;
        const uint32_t d = live ? ((uint32_t)(key >> shift) & mask) : 256u;       // dead lanes form their own group
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t below = __popc(peers & ((1u << lane) - 1u));
        if (live && below == 0) wc[warp][d] = (uint32_t)__popc(peers);
        __syncthreads();
        if (live) {
            uint32_t pre = 0;
            for (int w = 0; w < warp; ++w) pre += wc[w][d];
            out[base[d] + pre + below] = key;
        }
        __syncthreads();
        uint32_t tot = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { tot += wc[w][threadIdx.x]; wc[w][threadIdx.x] = 0; }
        base[threadIdx.x] += tot;
        __syncthreads();
    }
}

struct Sorter {
    uint32_t* hist;      // [256 * max_tiles + 1]
    uint32_t* tmp;       // scan scratch
    uint32_t* total;     // [1] scratch
    cudaStream_t st;
    // stable sort of a[0..n) on bits [lo, hi); b is the ping-pong buffer; returns the buffer that holds the result
    uint64_t* sort(uint64_t* a, uint64_t* b, uint32_t n, int lo, int hi) const {
        if (n < 2) return a;
        const uint32_t n_tiles = (n + RS_TILE - 1) / RS_TILE;
        for (int shift = lo; shift < hi; shift += 8) {
            const int width = hi - shift < 8 ? hi - shift : 8;
            const uint32_t mask = (1u << width) - 1u;
            { SPL_LAUNCH; k_rs_hist<<<n_tiles, RS_THREADS, 0, st>>>(a, n, shift, mask, hist, n_tiles); }
            launch_exscan_u32(hist, 256u * n_tiles, tmp, total, st);
            { SPL_LAUNCH; k_rs_scatter<<<n_tiles, RS_THREADS, 0, st>>>(a, b, n, shift, mask, hist, n_tiles); }
            uint64_t* t = a; a = b; b = t;
        }
        return a;
    }
};

__host__ __device__ inline int bits_for_u64(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }

// ---- phase A: endpoints -> sites ----------------------------------------------------------------
__global__ void k_gb_keys_a(const int32_t* __restrict__ jc, const int32_t* __restrict__ jl, const int32_t* __restrict__ jr,
                            const uint8_t* __restrict__ js, uint32_t J, int stranded, int pb, int vb, uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J) return;
    const uint64_t sb = (stranded && js[i] == '-') ? 1u : 0u;
    const uint64_t c = (uint64_t)(uint32_t)jc[i] << (pb + 1);
    keys[2 * i] = ((c | ((uint64_t)(uint32_t)jl[i] << 1) | sb) << vb) | (uint64_t)(2 * i);
    keys[2 * i + 1] = ((c | ((uint64_t)(uint32_t)jr[i] << 1) | sb) << vb) | (uint64_t)(2 * i + 1);
}

// flag[e] = 1 when sorted element e starts a new run of (key >> low)
__global__ void k_gb_heads(const uint64_t* __restrict__ k, uint32_t n, int low, uint32_t* __restrict__ flag) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    flag[e] = (e == 0 || (k[e] >> low) != (k[e - 1] >> low)) ? 1u : 0u;
}

__global__ void k_gb_sites(const uint64_t* __restrict__ k, const uint32_t* __restrict__ ex, uint32_t n, int vb, int pb, int stranded,
                           const uint8_t* __restrict__ js, GraphDev g, uint32_t* __restrict__ site_of) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t key = k[e] >> vb;
    const bool head = e == 0 || key != (k[e - 1] >> vb);
    const uint32_t idx = ex[e] - (head ? 0u : 1u);
    const uint32_t eid = (uint32_t)(k[e] & ((1ull << vb) - 1ull));
    site_of[eid] = idx;
    g.inc_line[e] = (int32_t)(eid >> 1);
    if (head) {
        const uint32_t line = eid >> 1;
        const int32_t chrom = (int32_t)(key >> (pb + 1));
        const uint8_t st = js[line];
        g.site_chrom[idx] = chrom;
        g.site_pos[idx] = (int32_t)((key >> 1) & ((1ull << pb) - 1ull));
        g.site_strand[idx] = st;                                       // first-seen strand (stable sort keeps row order)
        g.first_line[idx] = (int64_t)line;
        g.site_cls[idx] = !stranded ? 0 : st == '+' ? 1 : st == '-' ? 2 : 3;
        g.inc_off[idx] = (int32_t)e;
    }
    if (e == n - 1) g.inc_off[idx + 1] = (int32_t)n;
}

// per-chromosome site ranges: the table is sorted by chromosome, so cs_off[c] is a lower bound (one thread per chromosome)
__global__ void k_gb_cs_off(GraphDev g, int n_chrom, const uint32_t* __restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chrom) return;
    int lo = 0, hi = (int)counts[0];
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (g.site_chrom[mid] < c) lo = mid + 1; else hi = mid; }
    g.cs_off[c] = lo;
}
// layout of the direct-address bin index (one thread: a running sum over the chromosomes)
__global__ void k_gb_chrom_layout(int n_chrom, GraphDev g, uint32_t* __restrict__ counts) {
    if (blockIdx.x || threadIdx.x) return;
    int32_t nb = 0;
    for (int c = 0; c < n_chrom; ++c) {
        const int32_t s0 = g.cs_off[c], s1 = g.cs_off[c + 1];
        g.sb_base[c] = nb;
        nb += (s1 > s0 ? (max(g.site_pos[s1 - 1], 0) >> SB_SHIFT) + 1 : 0) + 1;     // + sentinel
    }
    g.sb_base[n_chrom] = nb;
    counts[1] = (uint32_t)nb;
}

__global__ void k_gb_sb_fill(GraphDev g, int n_chrom, uint32_t S, uint32_t n_entries) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries + 64u) return;
    if (i >= n_entries) { g.sb_off[i] = (int32_t)S; return; }
    int lo = 0, hi = n_chrom;                                           // last chromosome with sb_base <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((uint32_t)g.sb_base[mid] <= i) lo = mid; else hi = mid; }
    const int c = lo;
    const int32_t b = (int32_t)i - g.sb_base[c], nb = g.sb_base[c + 1] - g.sb_base[c] - 1;
    int s0 = g.cs_off[c], s1 = g.cs_off[c + 1];
    if (b >= nb) { g.sb_off[i] = s1; return; }
    const int32_t key = b << SB_SHIFT;
    while (s0 < s1) { const int mid = (s0 + s1) >> 1; if (g.site_pos[mid] < key) s0 = mid + 1; else s1 = mid; }
    g.sb_off[i] = s0;
}

// ---- phase B: directed edges -> PartnerCounts entries -------------------------------------------
__global__ void k_gb_keys_b(const uint32_t* __restrict__ site_of, uint32_t J, int sbits, int lb, uint64_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= J) return;
    const uint64_t a = site_of[2 * i], b = site_of[2 * i + 1];
    keys[2 * i] = (((a << sbits) | b) << lb) | i;                        // row order == first-appearance order
    keys[2 * i + 1] = (((b << sbits) | a) << lb) | i;
}

__global__ void k_gb_edges(const uint64_t* __restrict__ k, const uint32_t* __restrict__ ex, uint32_t n, int lb, int sbits, GraphDev g,
                           uint32_t* __restrict__ u_src, uint32_t* __restrict__ u_dst, uint32_t* __restrict__ u_first, uint32_t* __restrict__ u_lo) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint64_t pair = k[e] >> lb;
    const bool head = e == 0 || pair != (k[e - 1] >> lb);
    const uint32_t u = ex[e] - (head ? 0u : 1u);
    const uint32_t line = (uint32_t)(k[e] & ((1ull << lb) - 1ull));
    g.einc_line[e] = (int32_t)line;
    if (head) {
        u_dst[u] = (uint32_t)(pair & ((1ull << sbits) - 1ull));
        u_src[u] = (uint32_t)(pair >> sbits);
        u_first[u] = line;
        u_lo[u] = e;
    }
    if (e == n - 1) u_lo[u + 1] = n;
}

// ---- phase C: entries of a site in first-appearance order ---------------------------------------
__global__ void k_gb_keys_c(const uint32_t* __restrict__ u_src, const uint32_t* __restrict__ u_first, uint32_t E, int lb, int eb,
                            uint64_t* __restrict__ keys) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= E) return;
    keys[u] = ((((uint64_t)u_src[u] << lb) | u_first[u]) << eb) | u;
}

__global__ void k_gb_csr(const uint64_t* __restrict__ k, uint32_t E, uint32_t S, int eb, int lb, const uint32_t* __restrict__ u_dst,
                         const uint32_t* __restrict__ u_lo, GraphDev g, uint32_t* __restrict__ e_src) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= E) return;
    const uint32_t u = (uint32_t)(k[x] & ((1ull << eb) - 1ull));
    const uint32_t src = (uint32_t)(k[x] >> (eb + lb));
    const uint32_t dst = u_dst[u];
    g.pt_site[x] = (int32_t)dst;
    g.pc_pos[x] = g.site_pos[dst];
    g.einc_beg[x] = (int32_t)u_lo[u];
    g.einc_end[x] = (int32_t)u_lo[u + 1];
    e_src[x] = src;
    if (x == 0 || src != (uint32_t)(k[x - 1] >> (eb + lb))) g.pt_off[src] = (int32_t)x;      // every site has >= 1 entry
    if (x == E - 1) g.pt_off[S] = (int32_t)E;
}

// ---- phase D: competitors (S:364-372) ------------------------------------------------------------
__global__ void k_gb_cand_count(GraphDev g, uint32_t S, uint32_t* __restrict__ ncand) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    uint32_t n = 0;
    for (int a = g.pt_off[t]; a < g.pt_off[t + 1]; ++a) {
        const int p = g.pt_site[a];
        n += (uint32_t)(g.pt_off[p + 1] - g.pt_off[p]);
    }
    ncand[t] = n;
}

__global__ void k_gb_cand_fill(GraphDev g, uint32_t S, const uint32_t* __restrict__ cand_off, int pb, uint64_t* __restrict__ keys) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= S) return;
    uint32_t w = cand_off[t];
    const int32_t tp = g.site_pos[t];
    for (int a = g.pt_off[t]; a < g.pt_off[t + 1]; ++a) {
        const int p = g.pt_site[a];
        for (int q = g.pt_off[p]; q < g.pt_off[p + 1]; ++q) {
            const int32_t cpos = g.pc_pos[q];                            // position of the partner's partner
            keys[w++] = cpos != tp ? (((uint64_t)t << pb) | (uint32_t)cpos) : ((uint64_t)S << pb);   // own position: parked behind every site
        }
    }
}

__global__ void k_gb_cand_heads(const uint64_t* __restrict__ k, uint32_t n, int pb, uint32_t S, uint32_t* __restrict__ flag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = ((uint32_t)(k[i] >> pb) != S && (i == 0 || k[i] != k[i - 1])) ? 1u : 0u;
}

__global__ void k_gb_comp(const uint64_t* __restrict__ k, const uint32_t* __restrict__ ex, uint32_t n, int pb, uint32_t S, GraphDev g,
                          uint32_t* __restrict__ cp_cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = (uint32_t)(k[i] >> pb);
    if (t == S || !(i == 0 || k[i] != k[i - 1])) return;
    g.cp_pos[ex[i]] = (int32_t)(k[i] & ((1ull << pb) - 1ull));          // sorted by (site, position): compaction keeps the order
    atomicAdd(cp_cnt + t, 1u);
}

// ---- phase E: reverse-partner index ---------------------------------------------------------------
// the list hangs off the FIRST site index sharing (chromosome, position) with the partner; the clean regime has at
// most two sites per position ('+' then '-')
__device__ __forceinline__ uint32_t anchor_of(const GraphDev& g, uint32_t p) {
    return (p > 0 && g.site_pos[p - 1] == g.site_pos[p] && g.site_chrom[p - 1] == g.site_chrom[p]) ? p - 1 : p;
}
__global__ void k_gb_rp_count(GraphDev g, uint32_t E, uint32_t* __restrict__ rp_cnt) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= E) return;
    atomicAdd(rp_cnt + anchor_of(g, (uint32_t)g.pt_site[x]), 1u);
}
__global__ void k_gb_rp_fill(GraphDev g, uint32_t E, const uint32_t* __restrict__ e_src, uint32_t* __restrict__ rp_cur) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= E) return;
    const uint32_t a = anchor_of(g, (uint32_t)g.pt_site[x]);
    const uint32_t t = e_src[x];
    g.rp_site[(uint32_t)g.rp_off[a] + atomicAdd(rp_cur + a, 1u)] = (int32_t)t;      // order inside a list does not matter: only counted
    if (g.cp_off[t + 1] > g.cp_off[t]) g.site_hot[a] = 1;
}

__global__ void k_gb_widen(const int32_t* __restrict__ a, int64_t* __restrict__ o, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = a[i];
}
__global__ void k_gb_fill_i32(int32_t* a, uint32_t n, int32_t v) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

struct Carve {
    size_t off = 0;
    template <class T> size_t take(size_t n) {
        off = (off + 255) & ~(size_t)255;
        const size_t o = off;
        off += n * sizeof(T);
        return o;
    }
};

}  // namespace

bool graph_build_fits(int64_t J, int32_t n_chrom, int32_t max_pos) {
    if (J <= 0 || J >= ((int64_t)1 << 27)) return false;
    const int pb = bits_for_u64((uint64_t)(max_pos > 0 ? max_pos : 1)), cb = bits_for_u64((uint64_t)(n_chrom > 1 ? n_chrom - 1 : 1));
    const int vb = bits_for_u64((uint64_t)(2 * J - 1)), lb = bits_for_u64((uint64_t)(J - 1));
    const int sb = bits_for_u64((uint64_t)(2 * J));                     // S <= 2J (S itself is a sentinel value), E <= 2J
    return cb + pb + 1 + vb <= 64 && 2 * sb + lb <= 64 && sb + lb + vb <= 64 && pb + sb + 1 <= 64;
}

#define GB_CU(call)                                                                     \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return false; } \
    } while (0)

bool graph_build_device(GraphBuildMem& m, const int32_t* j_chrom, const int32_t* j_left, const int32_t* j_right, const uint8_t* j_strand,
                        const int64_t* j_score, int64_t n_junc, int32_t n_chrom, int32_t max_pos, bool stranded, void* stream,
                        int phase, GraphDev& g, GraphCounts& counts, std::string& err) {
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t J = (uint32_t)n_junc, n2 = 2 * J;
    const int pb = bits_for_u64((uint64_t)(max_pos > 0 ? max_pos : 1)), cb = bits_for_u64((uint64_t)(n_chrom > 1 ? n_chrom - 1 : 1));
    const int vb = bits_for_u64((uint64_t)(n2 - 1)), lb = bits_for_u64((uint64_t)(J - 1));

    // ---- allocations bounded by J (final arrays + workspace); cp_pos / sb_off follow once their sizes are known
    Carve f;
    const size_t o_cs = f.take<int32_t>((size_t)n_chrom + 1), o_sbb = f.take<int32_t>((size_t)n_chrom + 1);
    const size_t o_chrom = f.take<int32_t>(n2 + 1), o_pos = f.take<int32_t>(n2 + 64), o_strand = f.take<uint8_t>(n2 + 8), o_cls = f.take<uint8_t>(n2 + 8);
    const size_t o_hot = f.take<uint8_t>(n2 + 64), o_fl = f.take<int64_t>(n2 + 1);
    const size_t o_pto = f.take<int32_t>(n2 + 2), o_pts = f.take<int32_t>(n2 + 1), o_pcp = f.take<int32_t>(n2 + 1);
    const size_t o_cpo = f.take<int32_t>(n2 + 2), o_rpo = f.take<int32_t>(n2 + 2), o_rps = f.take<int32_t>(n2 + 1);
    const size_t o_ino = f.take<int32_t>(n2 + 2), o_inl = f.take<int32_t>(n2 + 1);
    const size_t o_eib = f.take<int32_t>(n2 + 1), o_eie = f.take<int32_t>(n2 + 1), o_eil = f.take<int32_t>(n2 + 1);
    const size_t o_js = f.take<int64_t>((size_t)J + 1);
    const size_t o_pto64 = f.take<int64_t>(n2 + 2), o_cpo64 = f.take<int64_t>(n2 + 2);
    const size_t o_jc = f.take<int32_t>(J), o_jl = f.take<int32_t>(J), o_jr = f.take<int32_t>(J), o_jst = f.take<uint8_t>((size_t)J + 8);
    GB_CU(m.fin.reserve(f.off + 256));
    char* fb = (char*)m.fin.p;
    g = GraphDev{};
    g.cs_off = (int32_t*)(fb + o_cs); g.sb_base = (int32_t*)(fb + o_sbb);
    g.site_chrom = (int32_t*)(fb + o_chrom); g.site_pos = (int32_t*)(fb + o_pos); g.site_strand = (uint8_t*)(fb + o_strand);
    g.site_cls = (uint8_t*)(fb + o_cls); g.site_hot = (uint8_t*)(fb + o_hot); g.first_line = (int64_t*)(fb + o_fl);
    g.pt_off = (int32_t*)(fb + o_pto); g.pt_site = (int32_t*)(fb + o_pts); g.pc_pos = (int32_t*)(fb + o_pcp);
    g.cp_off = (int32_t*)(fb + o_cpo); g.rp_off = (int32_t*)(fb + o_rpo); g.rp_site = (int32_t*)(fb + o_rps);
    g.inc_off = (int32_t*)(fb + o_ino); g.inc_line = (int32_t*)(fb + o_inl);
    g.einc_beg = (int32_t*)(fb + o_eib); g.einc_end = (int32_t*)(fb + o_eie); g.einc_line = (int32_t*)(fb + o_eil);
    g.j_score = (int64_t*)(fb + o_js); g.pt_off64 = (int64_t*)(fb + o_pto64); g.cp_off64 = (int64_t*)(fb + o_cpo64);
    int32_t* d_jc = (int32_t*)(fb + o_jc); int32_t* d_jl = (int32_t*)(fb + o_jl); int32_t* d_jr = (int32_t*)(fb + o_jr);
    uint8_t* d_js = (uint8_t*)(fb + o_jst);

    const uint32_t max_tiles = cdiv(n2, RS_TILE) + 1;
    Carve w;
    const size_t w_ka = w.take<uint64_t>(n2 + 2), w_kb = w.take<uint64_t>(n2 + 2), w_flag = w.take<uint32_t>(n2 + 2);
    const size_t w_hist = w.take<uint32_t>(256 * (size_t)max_tiles + 2), w_tmp = w.take<uint32_t>(exscan_tmp_words(256u * max_tiles) + exscan_tmp_words(n2 + 2) + 8);
    const size_t w_cnt = w.take<uint32_t>(16), w_site_of = w.take<uint32_t>(n2 + 2);
    const size_t w_usrc = w.take<uint32_t>(n2 + 2), w_udst = w.take<uint32_t>(n2 + 2), w_ufirst = w.take<uint32_t>(n2 + 2), w_ulo = w.take<uint32_t>(n2 + 2);
    const size_t w_esrc = w.take<uint32_t>(n2 + 2), w_nc = w.take<uint32_t>(n2 + 2), w_c1 = w.take<uint32_t>(n2 + 2), w_c2 = w.take<uint32_t>(n2 + 2);
    GB_CU(m.work.reserve(w.off + 256));
    char* wb = (char*)m.work.p;
    uint64_t* ka = (uint64_t*)(wb + w_ka); uint64_t* kb = (uint64_t*)(wb + w_kb);
    uint32_t* flag = (uint32_t*)(wb + w_flag);
    uint32_t* d_cnt = (uint32_t*)(wb + w_cnt);                            // [0] S, [1] bin entries, [2] E, [3] candidates, [4] C, [5] scratch
    uint32_t* site_of = (uint32_t*)(wb + w_site_of);
    uint32_t* u_src = (uint32_t*)(wb + w_usrc); uint32_t* u_dst = (uint32_t*)(wb + w_udst); uint32_t* u_first = (uint32_t*)(wb + w_ufirst);
    uint32_t* u_lo = (uint32_t*)(wb + w_ulo); uint32_t* e_src = (uint32_t*)(wb + w_esrc);
    uint32_t* ncand = (uint32_t*)(wb + w_nc); uint32_t* c1 = (uint32_t*)(wb + w_c1); uint32_t* c2 = (uint32_t*)(wb + w_c2);
    uint32_t* stmp = (uint32_t*)(wb + w_tmp) + exscan_tmp_words(256u * max_tiles) + 4;      // scan scratch of the phase kernels
    Sorter sorter{(uint32_t*)(wb + w_hist), (uint32_t*)(wb + w_tmp), d_cnt + 5, st};

    // ---- junction table to the device (phase 0: queued before the caller starts the big record upload, which
    // would otherwise sit in front of it on the host->device copy engine)
    if (phase == 0) {
    GB_CU(cudaMemcpyAsync(d_jc, j_chrom, (size_t)J * 4, cudaMemcpyHostToDevice, st));
    GB_CU(cudaMemcpyAsync(d_jl, j_left, (size_t)J * 4, cudaMemcpyHostToDevice, st));
    GB_CU(cudaMemcpyAsync(d_jr, j_right, (size_t)J * 4, cudaMemcpyHostToDevice, st));
    GB_CU(cudaMemcpyAsync(d_js, j_strand, (size_t)J, cudaMemcpyHostToDevice, st));
    GB_CU(cudaMemcpyAsync((void*)g.j_score, j_score, (size_t)J * 8, cudaMemcpyHostToDevice, st));
    counts.h2d_bytes = (double)J * 21.0;
    GB_CU(cudaMemsetAsync(d_cnt, 0, 64, st));
    return true;
    }

    // ---- A: sites
    { SPL_LAUNCH; k_gb_keys_a<<<cdiv(J, 256), 256, 0, st>>>(d_jc, d_jl, d_jr, d_js, J, stranded ? 1 : 0, pb, vb, ka); }
    uint64_t* sa = sorter.sort(ka, kb, n2, vb, vb + 1 + pb + cb);
    uint64_t* other = sa == ka ? kb : ka;
    { SPL_LAUNCH; k_gb_heads<<<cdiv(n2, 256), 256, 0, st>>>(sa, n2, vb, flag); }
    launch_exscan_u32(flag, n2, stmp, d_cnt + 0, st);
    { SPL_LAUNCH; k_gb_sites<<<cdiv(n2, 256), 256, 0, st>>>(sa, flag, n2, vb, pb, stranded ? 1 : 0, d_js, g, site_of); }
    { SPL_LAUNCH; k_gb_cs_off<<<cdiv((uint32_t)n_chrom + 1, 128), 128, 0, st>>>(g, n_chrom, d_cnt); }
    { SPL_LAUNCH; k_gb_chrom_layout<<<1, 32, 0, st>>>(n_chrom, g, d_cnt); }
    GB_CU(cudaGetLastError());
    GB_CU(cudaMemcpyAsync(m.h_cnt, d_cnt, 64, cudaMemcpyDeviceToHost, st));
    GB_CU(cudaStreamSynchronize(st));
    const uint32_t S = m.h_cnt[0], NB = m.h_cnt[1];
    if (S == 0 || S > n2) { err = "graph build: bad site count"; return false; }
    const int sbits = bits_for_u64((uint64_t)S);                          // S itself is used as a sentinel in phase D
    { SPL_LAUNCH; k_gb_fill_i32<<<1, 64, 0, st>>>(g.site_pos + S, 64, INT_MAX); }        // tail padding (never matches)
    GB_CU(cudaMemsetAsync(g.site_hot, 0, (size_t)S + 64, st));

    // ---- B: PartnerCounts entries
    { SPL_LAUNCH; k_gb_keys_b<<<cdiv(J, 256), 256, 0, st>>>(site_of, J, sbits, lb, other); }
    uint64_t* sbk = sorter.sort(other, sa, n2, lb, lb + 2 * sbits);
    other = sbk == ka ? kb : ka;
    { SPL_LAUNCH; k_gb_heads<<<cdiv(n2, 256), 256, 0, st>>>(sbk, n2, lb, flag); }
    launch_exscan_u32(flag, n2, stmp, d_cnt + 2, st);
    { SPL_LAUNCH; k_gb_edges<<<cdiv(n2, 256), 256, 0, st>>>(sbk, flag, n2, lb, sbits, g, u_src, u_dst, u_first, u_lo); }
    GB_CU(cudaGetLastError());
    GB_CU(cudaMemcpyAsync(m.h_cnt, d_cnt, 64, cudaMemcpyDeviceToHost, st));
    GB_CU(cudaStreamSynchronize(st));
    const uint32_t E = m.h_cnt[2];
    if (E == 0 || E > n2) { err = "graph build: bad edge count"; return false; }
    const int eb = bits_for_u64((uint64_t)(E - 1));

    // ---- C: first-appearance order inside a site
    { SPL_LAUNCH; k_gb_keys_c<<<cdiv(E, 256), 256, 0, st>>>(u_src, u_first, E, lb, eb, other); }
    uint64_t* sc = sorter.sort(other, other == ka ? kb : ka, E, eb, eb + lb + sbits);
    { SPL_LAUNCH; k_gb_csr<<<cdiv(E, 256), 256, 0, st>>>(sc, E, S, eb, lb, u_dst, u_lo, g, e_src); }

    // ---- D: competitors
    { SPL_LAUNCH; k_gb_cand_count<<<cdiv(S, 256), 256, 0, st>>>(g, S, ncand); }
    launch_exscan_u32(ncand, S, stmp, d_cnt + 3, st);
    GB_CU(cudaGetLastError());
    GB_CU(cudaMemcpyAsync(m.h_cnt, d_cnt, 64, cudaMemcpyDeviceToHost, st));
    GB_CU(cudaStreamSynchronize(st));
    const uint32_t NCAND = m.h_cnt[3];
    if ((uint64_t)NCAND >= ((uint64_t)1 << 31)) { err = "graph build: competitor candidate list exceeds 2^31"; return false; }
    {
        const uint32_t tiles = cdiv(NCAND, RS_TILE) + 1;
        Carve c;
        const size_t c_a = c.take<uint64_t>((size_t)NCAND + 2), c_b = c.take<uint64_t>((size_t)NCAND + 2), c_f = c.take<uint32_t>((size_t)NCAND + 2);
        const size_t c_h = c.take<uint32_t>(256 * (size_t)tiles + 2), c_t = c.take<uint32_t>(exscan_tmp_words(256u * tiles) + exscan_tmp_words(NCAND + 2) + 8);
        GB_CU(m.work2.reserve(c.off + 256));
        char* cbp = (char*)m.work2.p;
        uint64_t* da = (uint64_t*)(cbp + c_a); uint64_t* db = (uint64_t*)(cbp + c_b);
        uint32_t* dflag = (uint32_t*)(cbp + c_f);
        uint32_t* dtmp = (uint32_t*)(cbp + c_t) + exscan_tmp_words(256u * tiles) + 4;
        Sorter ds{(uint32_t*)(cbp + c_h), (uint32_t*)(cbp + c_t), d_cnt + 5, st};
        { SPL_LAUNCH; k_gb_cand_fill<<<cdiv(S, 256), 256, 0, st>>>(g, S, ncand, pb, da); }
        uint64_t* sd = ds.sort(da, db, NCAND, 0, pb + sbits);
        { SPL_LAUNCH; k_gb_cand_heads<<<cdiv(NCAND, 256), 256, 0, st>>>(sd, NCAND, pb, S, dflag); }
        launch_exscan_u32(dflag, NCAND, dtmp, d_cnt + 4, st);
        GB_CU(cudaGetLastError());
        GB_CU(cudaMemcpyAsync(m.h_cnt, d_cnt, 64, cudaMemcpyDeviceToHost, st));
        GB_CU(cudaStreamSynchronize(st));
        const uint32_t C = m.h_cnt[4];
        Carve x;
        const size_t x_cp = x.take<int32_t>((size_t)C + 1), x_sb = x.take<int32_t>((size_t)NB + 64);
        GB_CU(m.fin2.reserve(x.off + 256));
        g.cp_pos = (int32_t*)((char*)m.fin2.p + x_cp); g.sb_off = (int32_t*)((char*)m.fin2.p + x_sb);
        GB_CU(cudaMemsetAsync(c1, 0, ((size_t)S + 1) * 4, st));           // cp_cnt
        { SPL_LAUNCH; k_gb_comp<<<cdiv(NCAND, 256), 256, 0, st>>>(sd, dflag, NCAND, pb, S, g, c1); }
        GB_CU(cudaMemcpyAsync(g.cp_off, c1, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
        launch_exscan_u32((uint32_t*)g.cp_off, S, stmp, (uint32_t*)g.cp_off + S, st);
        counts.C = C;
    }

    // ---- E: reverse partners, hot flags, bin index
    GB_CU(cudaMemsetAsync(c1, 0, ((size_t)S + 1) * 4, st));               // rp_cnt
    GB_CU(cudaMemsetAsync(c2, 0, ((size_t)S + 1) * 4, st));               // rp_cur
    { SPL_LAUNCH; k_gb_rp_count<<<cdiv(E, 256), 256, 0, st>>>(g, E, c1); }
    GB_CU(cudaMemcpyAsync(g.rp_off, c1, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
    launch_exscan_u32((uint32_t*)g.rp_off, S, stmp, (uint32_t*)g.rp_off + S, st);
    { SPL_LAUNCH; k_gb_rp_fill<<<cdiv(E, 256), 256, 0, st>>>(g, E, e_src, c2); }
    { SPL_LAUNCH; k_gb_sb_fill<<<cdiv(NB + 64, 256), 256, 0, st>>>(g, n_chrom, S, NB); }
    { SPL_LAUNCH; k_gb_widen<<<cdiv(S + 1, 256), 256, 0, st>>>(g.pt_off, g.pt_off64, S + 1); }
    { SPL_LAUNCH; k_gb_widen<<<cdiv(S + 1, 256), 256, 0, st>>>(g.cp_off, g.cp_off64, S + 1); }
    GB_CU(cudaGetLastError());
    counts.S = S; counts.E = E; counts.NB = NB;
    return true;
}

}  // namespace spl
