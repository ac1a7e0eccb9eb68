/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the counting path of SpliSER v0.1.8, used as
 * (a) the checker of the CUDA path at sizes the pure-Python oracle cannot reach and (b) the
 * "port" CPU baseline that bench.py times on the GPU box's host cores.  The product never links,
 * loads or calls this file.
 *
 * It keeps the reference's algorithm: a per-chromosome list of Site records built line by line
 * with the hand-rolled bisection (SpliSER_v0_1_8.py:175-225, :289-355), then FOR EACH SITE a
 * fetch of the reads overlapping chr:t-(t+1) (S:422, restated with htslib's overlap rule because
 * samtools is not installed here) and the per-CIGAR-operator state machine of checkBam
 * (S:427-559), then findBeta2Counts (S:581-623) and calculateSSE (S:626-639).  The only liberty
 * is the read fetch: records are indexed per chromosome by start position (what the .bai does
 * for samtools) instead of forking a process per site, and sites are processed in parallel with
 * OpenMP (each site is independent in the reference as well).
 *
 * Parity: checked against oracle/spliser_oracle.py -- itself checked against the unmodified
 * reference -- by tests/test_oracle.py on every golden fixture and on seeded fuzz.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

#define FLAG_STRANDED 1u
#define FLAG_RF 2u
#define FLAG_CRYPTIC 4u
#define FLAG_COMBINE 8u

typedef struct { int32_t pos; int64_t cnt; } PosCnt;

typedef struct Site {
    int32_t chrom, pos;
    uint8_t strand;
    int64_t first_line;
    int64_t alpha, beta1, beta2s, beta2c;
    double beta2w, sse;
    struct Site** partners; int n_partners, cap_partners;      /* Site.Partners, G:260-262 */
    PosCnt* pcounts; int n_pcounts, cap_pcounts;               /* Site.PartnerCounts, G:243-246 */
    PosCnt* dc; int n_dc, cap_dc;                              /* PartnerBeta2DoubleCounts, G:249-258 */
    int32_t* comp; int n_comp, cap_comp;                       /* CompetitorPos, G:264-266 */
} Site;

typedef struct { Site** v; int64_t n, cap; } SiteList;

/* ---- records view (same layout as spl_records_view of include/spliser_b200.h) ---------------- */
typedef struct {
    int64_t n_rec, n_cigar;
    const int32_t* pos; const uint16_t* flag; const uint32_t* cig_off; const uint32_t* cigar;
    int32_t n_seg; const int32_t* seg_chrom; const int64_t* seg_off;
} RecView;

typedef struct {
    int64_t n_sites;
    int32_t* chrom; int32_t* pos; uint8_t* strand; int64_t* first_line;
    int64_t* alpha; int64_t* beta1; int64_t* beta2s; int64_t* beta2c; double* beta2w; double* sse;
    int64_t* pc_off; int32_t* pc_pos; int64_t* pc_cnt;
    int64_t* cp_off; int32_t* cp_pos;
} OracleResult;

static int is_pm(uint8_t s) { return s == '+' || s == '-'; }

static void* grow(void* p, int* cap, int need, size_t es) {
    if (need <= *cap) return p;
    int c = *cap ? *cap * 2 : 4;
    while (c < need) c *= 2;
    *cap = c;
    return realloc(p, (size_t)c * es);
}

static PosCnt* pc_find(PosCnt* a, int n, int32_t pos) {
    for (int i = 0; i < n; ++i) if (a[i].pos == pos) return &a[i];
    return NULL;
}
static void pc_add(PosCnt** a, int* n, int* cap, int32_t pos, int64_t cnt) {
    PosCnt* e = pc_find(*a, *n, pos);
    if (!e) {
        *a = (PosCnt*)grow(*a, cap, *n + 1, sizeof(PosCnt));
        e = &(*a)[(*n)++];
        e->pos = pos; e->cnt = 0;
    }
    e->cnt += cnt;
}

/* Site.addCompetitorPos, G:264-266: sorted unique */
static void comp_add(Site* s, int32_t pos) {
    int i = 0;
    while (i < s->n_comp && s->comp[i] < pos) ++i;
    if (i < s->n_comp && s->comp[i] == pos) return;
    s->comp = (int32_t*)grow(s->comp, &s->cap_comp, s->n_comp + 1, sizeof(int32_t));
    memmove(s->comp + i + 1, s->comp + i, (size_t)(s->n_comp - i) * sizeof(int32_t));
    s->comp[i] = pos;
    s->n_comp++;
}
static int comp_has(const Site* s, int32_t pos) {
    for (int i = 0; i < s->n_comp; ++i) if (s->comp[i] == pos) return 1;
    return 0;
}

/* Site.__eq__ / __lt__, G:123-163 (a fall-through None is falsy) */
static int site_eq(const Site* a, const Site* b, int stranded) {
    if (stranded) return a->pos == b->pos && a->strand == b->strand;
    return a->pos == b->pos;
}
static int site_lt(const Site* a, const Site* b, int stranded) {
    if (stranded && a->pos == b->pos) {
        if (a->strand == b->strand) return 0;
        return a->strand == '+' && b->strand == '-';
    }
    return a->pos < b->pos;
}

static int strand_ok_q(uint8_t q, uint8_t s, int stranded) { return q == s || !stranded || !is_pm(q); }

/* binary_site_search, S:175-225 */
static int64_t site_search(const SiteList* L, int32_t pos, uint8_t strand, int stranded) {
    const int64_t length = L->n;
    if (length == 0) return -1;
    int64_t idx = length / 2, past_max = length, past_min = 0, last_idx = -1, new_idx = idx;
    int stuck = 0, found = 0;
    while (!stuck && !found) {
        const int32_t p = L->v[idx]->pos;
        if (pos == p) {
            if (strand_ok_q(strand, L->v[idx]->strand, stranded)) { found = 1; break; }
            const int64_t cand[2] = {idx - 1, idx + 1};
            for (int k = 0; k < 2; ++k) {
                const int64_t a = cand[k];
                if (a >= 0 && a < length && pos == L->v[a]->pos && strand_ok_q(strand, L->v[a]->strand, stranded)) { idx = a; found = 1; }
            }
            break;
        } else if (pos >= p) {
            new_idx = idx + ((past_max - idx) / 2); past_min = idx;
        } else {
            new_idx = idx - ((idx - past_min) / 2); past_max = idx;
            if (idx == 1) new_idx = 0;
        }
        if (idx != last_idx) { last_idx = idx; idx = new_idx; } else stuck = 1;
    }
    return found ? idx : -1;
}

static void insort(SiteList* L, Site* x, int stranded) {      /* bisect.insort_right, S:347-350 */
    int64_t lo = 0, hi = L->n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (site_lt(x, L->v[mid], stranded)) hi = mid; else lo = mid + 1;
    }
    if (L->n + 1 > L->cap) {
        L->cap = L->cap ? L->cap * 2 : 64;
        L->v = (Site**)realloc(L->v, (size_t)L->cap * sizeof(Site*));
    }
    memmove(L->v + lo + 1, L->v + lo, (size_t)(L->n - lo) * sizeof(Site*));
    L->v[lo] = x;
    L->n++;
}

static void link_sites(Site* a, Site* b, int64_t alpha, int stranded) {   /* S:352-355 */
    int present = 0;
    for (int i = 0; i < a->n_partners && !present; ++i) present = (a->partners[i] == b) || site_eq(a->partners[i], b, stranded);
    if (!present) {
        a->partners = (Site**)grow(a->partners, &a->cap_partners, a->n_partners + 1, sizeof(Site*));
        a->partners[a->n_partners++] = b;
    }
    pc_add(&a->pcounts, &a->n_pcounts, &a->cap_pcounts, b->pos, alpha);
}

/* findAlphaCounts S:289-355 + findCompetitorPos S:364-372 */
static SiteList* build_sites(int32_t n_chrom, int64_t n_junc, const int32_t* jc, const int32_t* jl, const int32_t* jr,
                             const int64_t* js, const uint8_t* jst, int stranded) {
    SiteList* lists = (SiteList*)calloc((size_t)(n_chrom > 0 ? n_chrom : 1), sizeof(SiteList));
    for (int64_t i = 0; i < n_junc; ++i) {
        SiteList* L = &lists[jc[i]];
        const int64_t li = site_search(L, jl[i], jst[i], stranded), ri = site_search(L, jr[i], jst[i], stranded);
        Site* pair[2];
        const int64_t idx[2] = {li, ri};
        const int32_t pp[2] = {jl[i], jr[i]};
        for (int k = 0; k < 2; ++k) {
            Site* s;
            if (idx[k] < 0) {
                s = (Site*)calloc(1, sizeof(Site));
                s->chrom = jc[i]; s->pos = pp[k]; s->strand = jst[i]; s->first_line = i;
            } else s = L->v[idx[k]];
            s->alpha += js[i];
            pair[k] = s;
        }
        if (li < 0) insort(L, pair[0], stranded);
        if (ri < 0) insort(L, pair[1], stranded);
        link_sites(pair[0], pair[1], js[i], stranded);
        link_sites(pair[1], pair[0], js[i], stranded);
    }
    for (int32_t c = 0; c < n_chrom; ++c)
        for (int64_t k = 0; k < lists[c].n; ++k) {
            Site* s = lists[c].v[k];
            for (int a = 0; a < s->n_partners; ++a) {
                Site* p = s->partners[a];
                for (int b = 0; b < p->n_partners; ++b)
                    if (p->partners[b]->pos != s->pos) comp_add(s, p->partners[b]->pos);
            }
        }
    return lists;
}

/* check_strand, S:374-406 */
static int check_strand(uint32_t mode, uint32_t flag, uint8_t site_strand) {
    const int first = (flag & 64u) || !(flag & 1u);
    const int rev = (flag & 16u) != 0;
    uint8_t rs;
    if (!(mode & FLAG_RF)) rs = first ? (rev ? '-' : '+') : (rev ? '+' : '-');
    else rs = first ? (rev ? '+' : '-') : (rev ? '-' : '+');
    return rs == site_strand;
}

/* per-chromosome read index: record ids sorted by position + the largest reference span */
typedef struct { int64_t n; int64_t* id; int32_t* start; int32_t max_span; } ChromIndex;

static int cmp_pos(const void* a, const void* b, void* ctx) {
    const int32_t* pos = (const int32_t*)ctx;
    const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    if (pos[x] != pos[y]) return pos[x] < pos[y] ? -1 : 1;
    return x < y ? -1 : (x > y);
}

static int32_t ref_len(const RecView* r, int64_t i) {
    int64_t rl = 0;
    for (uint32_t k = r->cig_off[i]; k < r->cig_off[i + 1]; ++k) {
        const uint32_t op = r->cigar[k] & 15u;
        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += r->cigar[k] >> 4;
    }
    return (int32_t)rl;
}

static ChromIndex* build_index(const RecView* r, int32_t n_chrom) {
    ChromIndex* ix = (ChromIndex*)calloc((size_t)(n_chrom > 0 ? n_chrom : 1), sizeof(ChromIndex));
    for (int32_t s = 0; s < r->n_seg; ++s) {
        const int32_t c = r->seg_chrom[s];
        if (c >= 0 && c < n_chrom) ix[c].n += r->seg_off[s + 1] - r->seg_off[s];
    }
    for (int32_t c = 0; c < n_chrom; ++c) {
        ix[c].id = (int64_t*)malloc((size_t)(ix[c].n + 1) * sizeof(int64_t));
        ix[c].start = (int32_t*)malloc((size_t)(ix[c].n + 1) * sizeof(int32_t));
        ix[c].n = 0;
    }
    for (int32_t s = 0; s < r->n_seg; ++s) {
        const int32_t c = r->seg_chrom[s];
        if (c < 0 || c >= n_chrom) continue;
        for (int64_t i = r->seg_off[s]; i < r->seg_off[s + 1]; ++i) ix[c].id[ix[c].n++] = i;
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t c = 0; c < n_chrom; ++c) {
        int sorted = 1;
        for (int64_t k = 1; k < ix[c].n && sorted; ++k) sorted = r->pos[ix[c].id[k - 1]] <= r->pos[ix[c].id[k]];
        if (!sorted) qsort_r(ix[c].id, (size_t)ix[c].n, sizeof(int64_t), cmp_pos, (void*)r->pos);
        int32_t ms = 1;
        for (int64_t k = 0; k < ix[c].n; ++k) {
            ix[c].start[k] = r->pos[ix[c].id[k]];
            int32_t rl = ref_len(r, ix[c].id[k]);
            if (rl < 1) rl = 1;
            if (rl > ms) ms = rl;
        }
        ix[c].max_span = ms;
    }
    return ix;
}

/* checkBam, S:408-559, for one site */
static void check_bam(Site* site, const RecView* r, const ChromIndex* ix, uint32_t mode) {
    const int32_t t = site->pos;
    const int stranded = (mode & FLAG_STRANDED) != 0, combine = (mode & FLAG_COMBINE) != 0;
    /* records with POS <= t (S:435) whose reference span reaches t: start in [t - max_span, t] */
    int64_t lo = 0, hi = ix->n;
    while (lo < hi) { const int64_t mid = (lo + hi) / 2; if (ix->start[mid] <= t) lo = mid + 1; else hi = mid; }
    const int64_t last = lo;                 /* first index with start > t */
    int32_t sites_buf[64];
    for (int64_t q = last - 1; q >= 0 && ix->start[q] >= t - ix->max_span; --q) {
        const int64_t i = ix->id[q];
        const int32_t pos1 = r->pos[i];
        int32_t rl = ref_len(r, i);
        if (rl < 1) rl = 1;
        /* samtools view chr:t-(t+1): 0-based [t-1, t+1) overlaps [pos1-1, pos1-1+rl) */
        if (!(pos1 - 1 < t + 1 && pos1 - 1 + rl > t - 1)) continue;
        if (pos1 > t) continue;
        const uint32_t flag = r->flag[i];
        int32_t* ss = sites_buf; int n_ss = 0, cap_ss = 64;
        int32_t partner_used = 0; int have_pu = 0;
        int comp = 0, alpha_read = 0, beta1_read = 0, flanking = 0, mutex = 0;
        int32_t cpos = -1;
        int32_t cur = pos1;
        int mapped = 0, progression = 0;
        for (uint32_t k = r->cig_off[i]; k < r->cig_off[i + 1]; ++k) {
            const uint32_t op = r->cigar[k] & 15u;
            const int32_t d = (int32_t)(r->cigar[k] >> 4);
            if (op == 0 || op == 7 || op == 8) { mapped = 1; progression = 1; }       /* M = X, S:457-459 */
            else if (op == 3 || op == 2) { mapped = 0; progression = 1; }              /* N D,   S:460-462 */
            else if (op == 1 || op == 4 || op == 5 || op == 6) progression = 0;        /* I S H P, S:463-464 */
            if (!progression) continue;
            cur += d;
            if (t >= cur - d && cur > t && cur > t + 1) {                              /* S:469 */
                if (mapped) {
                    if (stranded) { if (check_strand(mode, flag, site->strand)) beta1_read = 1; }
                    else beta1_read = 1;
                }
            }
            if (op == 3) {
                const int32_t l = cur - d - 1, rr = cur - 1;                           /* S:482-483 */
                if (n_ss + 2 > cap_ss) {
                    cap_ss *= 2;
                    int32_t* nb = (int32_t*)malloc((size_t)cap_ss * sizeof(int32_t));
                    memcpy(nb, ss, (size_t)n_ss * sizeof(int32_t));
                    if (ss != sites_buf) free(ss);
                    ss = nb;
                }
                ss[n_ss++] = l; ss[n_ss++] = rr;
                if (l == t) { partner_used = rr; have_pu = 1; alpha_read = 1; }
                if (rr == t) { partner_used = l; have_pu = 1; alpha_read = 1; }
                if (comp_has(site, rr) && pc_find(site->pcounts, site->n_pcounts, l)) { comp = 1; cpos = rr; }
                if (comp_has(site, l) && pc_find(site->pcounts, site->n_pcounts, rr)) { comp = 1; cpos = l; }
                if (comp && t > l && t < rr) flanking = 1;                             /* S:503-505 */
                if (!alpha_read && !comp && t > l && t < rr) {                         /* S:507-512 */
                    if (stranded) { if (check_strand(mode, flag, site->strand)) mutex = 1; }
                    else mutex = 1;
                }
            }
        }
        const int beta1type = beta1_read && comp;                                      /* S:516-517 */
        if (alpha_read && comp) {                                                      /* S:519-527 */
            for (int e = 0; e < site->n_pcounts; ++e) {
                const int32_t p = site->pcounts[e].pos;
                int in = 0;
                for (int z = 0; z < n_ss && !in; ++z) in = ss[z] == p;
                if (in && !(have_pu && p == partner_used)) pc_add(&site->dc, &site->n_dc, &site->cap_dc, p, 1);
            }
        } else if (flanking) {                                                         /* S:529-535 */
            if (combine) { site->beta2s += 1; comp_add(site, cpos); }
        } else if (mutex) {                                                            /* S:540-541 */
            site->beta2s += 1;
        } else if (beta1type) {                                                        /* S:544-556 */
            for (int e = 0; e < site->n_pcounts; ++e) {
                const int32_t p = site->pcounts[e].pos;
                int in = 0;
                for (int z = 0; z < n_ss && !in; ++z) in = ss[z] == p;
                if (in) pc_add(&site->dc, &site->n_dc, &site->cap_dc, p, 1);
            }
            site->beta2s += 1;
            comp_add(site, cpos);
        } else if (beta1_read) {                                                       /* S:558-559 */
            site->beta1 += 1;
        }
        if (ss != sites_buf) free(ss);
    }
}

/* findBeta2Counts S:581-623 + calculateSSE S:626-639 */
static void beta2_and_sse(Site* s, uint32_t mode) {
    int64_t b2c = 0;
    double b2w = 0.0;
    for (int a = 0; a < s->n_partners; ++a) {
        Site* p = s->partners[a];
        for (int e = 0; e < p->n_pcounts; ++e) {
            const int32_t cp = p->pcounts[e].pos;
            if ((p->pos > s->pos && cp < s->pos) || (p->pos < s->pos && cp > s->pos)) {
                s->beta2s += p->pcounts[e].cnt;
                pc_add(&s->dc, &s->n_dc, &s->cap_dc, p->pos, p->pcounts[e].cnt);
            }
        }
        const PosCnt* pc = pc_find(s->pcounts, s->n_pcounts, p->pos);
        const int64_t pcount = pc ? pc->cnt : 0;
        int64_t b2 = p->alpha - pcount;
        const PosCnt* d = pc_find(s->dc, s->n_dc, p->pos);
        if (d) { b2 -= d->cnt; if (b2 < 0) b2 = 0; }
        b2c += b2;
        const double w = s->alpha > 0 ? (double)pcount / (double)s->alpha : 0.0;
        volatile double term = (double)b2 * w;             /* mul, then add: no fused contraction */
        b2w = b2w + term;
    }
    s->beta2c += b2c;
    s->beta2w = b2w;
    if (mode & FLAG_CRYPTIC) {
        const double betas = (double)(s->beta1 + s->beta2s) + s->beta2w;
        const double den = (double)s->alpha + betas;
        s->sse = den > 0.0 ? (double)s->alpha / den : 0.0;
    } else {
        const int64_t den = s->alpha + s->beta1 + s->beta2s;
        s->sse = den > 0 ? (double)s->alpha / (double)den : 0.0;
    }
}

static void free_site(Site* s) { free(s->partners); free(s->pcounts); free(s->dc); free(s->comp); free(s); }

static void free_index(ChromIndex* ix, int32_t n_chrom) {
    for (int32_t c = 0; c < n_chrom; ++c) { free(ix[c].id); free(ix[c].start); }
    free(ix);
}

/* ---- exported entry points (ctypes, see oracle/c_oracle.py) ----------------------------------- */
int oracle_process(const RecView* r, int32_t n_chrom, int64_t n_junc, const int32_t* jc, const int32_t* jl,
                   const int32_t* jr, const int64_t* js, const uint8_t* jst, uint32_t mode, int n_threads, OracleResult* out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    const int stranded = (mode & FLAG_STRANDED) != 0;
    const double t0 = now_s();
    SiteList* lists = build_sites(n_chrom, n_junc, jc, jl, jr, js, jst, stranded);
    const double t1 = now_s();
    ChromIndex* ix = build_index(r, n_chrom);
    const double t2 = now_s();
    int64_t S = 0;
    for (int32_t c = 0; c < n_chrom; ++c) S += lists[c].n;
    for (int32_t c = 0; c < n_chrom; ++c) {                                  /* processSites, S:681-692 */
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t k = 0; k < lists[c].n; ++k) check_bam(lists[c].v[k], r, &ix[c], mode);
        for (int64_t k = 0; k < lists[c].n; ++k) beta2_and_sse(lists[c].v[k], mode);
    }
    if (getenv("ORACLE_TIMING")) fprintf(stderr, "oracle: build_sites %.3fs build_index %.3fs count %.3fs\n", t1 - t0, t2 - t1, now_s() - t2);
    memset(out, 0, sizeof *out);
    out->n_sites = S;
    const size_t n1 = (size_t)S + 1;
    out->chrom = (int32_t*)calloc(n1, 4); out->pos = (int32_t*)calloc(n1, 4); out->strand = (uint8_t*)calloc(n1, 1);
    out->first_line = (int64_t*)calloc(n1, 8);
    out->alpha = (int64_t*)calloc(n1, 8); out->beta1 = (int64_t*)calloc(n1, 8); out->beta2s = (int64_t*)calloc(n1, 8);
    out->beta2c = (int64_t*)calloc(n1, 8); out->beta2w = (double*)calloc(n1, 8); out->sse = (double*)calloc(n1, 8);
    out->pc_off = (int64_t*)calloc(n1, 8); out->cp_off = (int64_t*)calloc(n1, 8);
    int64_t E = 0, CP = 0, k = 0;
    for (int32_t c = 0; c < n_chrom; ++c)
        for (int64_t q = 0; q < lists[c].n; ++q) { E += lists[c].v[q]->n_pcounts; CP += lists[c].v[q]->n_comp; }
    out->pc_pos = (int32_t*)calloc((size_t)E + 1, 4); out->pc_cnt = (int64_t*)calloc((size_t)E + 1, 8);
    out->cp_pos = (int32_t*)calloc((size_t)CP + 1, 4);
    E = CP = 0;
    for (int32_t c = 0; c < n_chrom; ++c)
        for (int64_t q = 0; q < lists[c].n; ++q, ++k) {
            Site* s = lists[c].v[q];
            out->chrom[k] = s->chrom; out->pos[k] = s->pos; out->strand[k] = s->strand; out->first_line[k] = s->first_line;
            out->alpha[k] = s->alpha; out->beta1[k] = s->beta1; out->beta2s[k] = s->beta2s; out->beta2c[k] = s->beta2c;
            out->beta2w[k] = s->beta2w; out->sse[k] = s->sse;
            for (int e = 0; e < s->n_pcounts; ++e) { out->pc_pos[E] = s->pcounts[e].pos; out->pc_cnt[E++] = s->pcounts[e].cnt; }
            for (int e = 0; e < s->n_comp; ++e) out->cp_pos[CP++] = s->comp[e];
            out->pc_off[k + 1] = E; out->cp_off[k + 1] = CP;
        }
    for (int32_t c = 0; c < n_chrom; ++c) {
        for (int64_t q = 0; q < lists[c].n; ++q) free_site(lists[c].v[q]);
        free(lists[c].v);
    }
    free(lists);
    free_index(ix, n_chrom);
    return 0;
}

void oracle_result_free(OracleResult* o) {
    free(o->chrom); free(o->pos); free(o->strand); free(o->first_line); free(o->alpha); free(o->beta1); free(o->beta2s);
    free(o->beta2c); free(o->beta2w); free(o->sse); free(o->pc_off); free(o->pc_pos); free(o->pc_cnt); free(o->cp_off);
    free(o->cp_pos);
    memset(o, 0, sizeof *o);
}

/* the checkBam call of combine (S:899-904) for a list of gap sites */
int oracle_recount(const RecView* r, int32_t n_chrom, int64_t n_sites, const int32_t* s_chrom, const int32_t* s_pos,
                   const uint8_t* s_strand, const int64_t* p_off, const int32_t* p_pos, const int64_t* c_off,
                   const int32_t* c_pos, uint32_t mode, int n_threads, int64_t* beta1_out, int64_t* beta2s_out) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    mode |= FLAG_COMBINE;
    ChromIndex* ix = build_index(r, n_chrom);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t i = 0; i < n_sites; ++i) {
        beta1_out[i] = 0; beta2s_out[i] = 0;
        if (s_chrom[i] < 0 || s_chrom[i] >= n_chrom) continue;
        Site* s = (Site*)calloc(1, sizeof(Site));
        s->chrom = s_chrom[i]; s->pos = s_pos[i]; s->strand = s_strand[i];
        for (int64_t e = p_off[i]; e < p_off[i + 1]; ++e) pc_add(&s->pcounts, &s->n_pcounts, &s->cap_pcounts, p_pos[e], 0);
        for (int64_t e = c_off[i]; e < c_off[i + 1]; ++e) comp_add(s, c_pos[e]);
        check_bam(s, r, &ix[s->chrom], mode);
        beta1_out[i] = s->beta1; beta2s_out[i] = s->beta2s;
        free_site(s);
    }
    free_index(ix, n_chrom);
    return 0;
}
