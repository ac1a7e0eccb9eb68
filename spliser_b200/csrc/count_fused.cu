// Fused counting kernel: alignment records resident in HBM -> per-site counters, one pass, nothing materialised.
//
// Replaces processSites -> checkBam per site (SpliSER_v0_1_8.py:681-692, :408-559) by one traversal of the records.
// This is the difference-array / prefix-scan formulation of the stabbing count: every advancing CIGAR operator [cur, cur+len)
// of type M/=/X (S:457-459, :469) or N (S:480-483, :507-512) stabs the sites with position in [cur, cur+len-2], i.e. the
// contiguous site index range [lb(cur), lb(cur+len-1)) of the sorted site table.  The range gets +1 / -1 in the class's
// difference array (cnt.diff, cov words for M, span words for N); k_finalize takes the prefix sums.  The stabbing
// variant (k_beta1_stab in kernels.cu, streaming a bin-partitioned copy of the blocks) is kept for the cross-check.
//
// Layout of the kernel: persistent CTAs (2 per SM), one producer warp + 16 consumer warps.
//   producer  claims chunks of <= FC_RECS records from a global counter and stages their pos / flag / cig_off / CIGAR
//             slices into a 3-stage shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier);
//   consumers one record per lane.  A lane walks its CIGAR from shared memory and carries the site index of its
//             current reference position along (one lookup through the direct-address bin index per read or long
//             jump, short walks over the L1-resident site table otherwise).  Lanes run their operators in lock step, so
//             that equal (index, class, kind) targets of neighbouring lanes -- coordinate-sorted reads hit the same
//             few sites -- merge into one RED per run.
//   A junction endpoint that sits on a "hot" site (an anchor whose reverse-partner list has competitors) may make
//   compSplicing true for some site (S:494-501): the (record, operator, side) goes to the warp's shared-memory list and
//   is classified after the warp's records of the chunk, all lanes busy, from the CIGAR still staged (S:503-557).
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <mutex>

#include "dev_helpers.cuh"
#include "device_types.h"

namespace spl {

namespace {

constexpr int FC_STAGES = 3;
constexpr int FC_CWARPS = 16;
constexpr int FC_CONSUMERS = FC_CWARPS * 32;
constexpr int FC_THREADS = FC_CONSUMERS + 32;
constexpr int FC_RPAD = 16;                 // slack for the 16-byte alignment of the staged record slices
constexpr int FC_CIG = 4096;                // staged CIGAR words per stage (a chunk with more reads them from global memory)
constexpr int FC_LIST = 128;                // hot items per warp list
constexpr int FC_BATCH = 2;                 // chunks a producer claims per atomic
constexpr int FC_WALK = 256;                // operators longer than this jump through the bin index instead of walking
constexpr int FC_MAXJ = 4, FC_MAXB = 6;     // junctions / blocks of a read kept in registers by the exception path

constexpr uint32_t FM_DONE = 1u, FM_GLOBAL_CIG = 2u;

struct FStage {
    int32_t  pos[FC_RECS + FC_RPAD];
    uint32_t off[FC_RECS + FC_RPAD];
    uint32_t cig[FC_CIG + 8];
    uint16_t flag[FC_RECS + FC_RPAD];
};
static_assert(sizeof(FStage) % 16 == 0 && ((FC_RECS + FC_RPAD) * 4) % 16 == 0 && ((FC_CIG + 8) * 4) % 16 == 0, "TMA destinations are 16-byte aligned");

struct FMeta {
    uint32_t flags, n_rec, skip, cig_base;  // skip: staged index of the chunk's first record; cig_base: absolute index of staged word 0
    int32_t  chrom, s0, s1, sb_g0, sb_nb;
};

struct FSmem {
    FStage st[FC_STAGES];
    unsigned long long list[FC_CWARPS][FC_LIST];
    FMeta meta[FC_STAGES];
    uint64_t full[FC_STAGES], empty[FC_STAGES];
};

struct FArgs {
    DevRecords rec;
    const FChunk* chunks;
    uint32_t chunk_lo, chunk_hi;
    DevGraph g;
    DevCounters cnt;
    uint32_t* work;
    uint32_t mode;
};

// +n on word `key` of the difference arrays for every run of neighbouring lanes with the same key (v: the lane takes part)
__device__ __forceinline__ void run_add(uint32_t* base, uint32_t key, bool v, bool neg, int lane) {
    const uint32_t kk = v ? key : 0xffffffffu;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, kk, 1);
    const bool head = lane == 0 || kk != prev;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (head && v) {
        const uint32_t after = heads & ~((2u << lane) - 1u);           // heads above this lane
        const uint32_t n = (uint32_t)((after ? __ffs(after) - 1 : 32) - lane);
        atomicAdd(base + key, neg ? 0u - n : n);
    }
}

// the staged (or, for oversized chunks, global) CIGAR words of a chunk
template <bool STAGED>
struct CigSrc {
    const uint32_t* p;       // STAGED: stage array, indexed by (absolute index - base); else the global array
    uint32_t base;
    __device__ __forceinline__ uint32_t operator()(uint32_t c) const { return STAGED ? p[c - base] : __ldg(p + c); }
};

// first site index >= the caller's lower bound `from` with position >= pos, through the chromosome's bin index
__device__ __forceinline__ int bin_lower(const DevGraph& g, const FMeta& m, int32_t pos, int from) {
    int i = __ldg(g.sb_off + m.sb_g0 + min(max(pos, 0) >> SB_SHIFT, m.sb_nb));
    i = max(i, from);
    while (i < m.s1 && __ldg(g.site_pos + i) < pos) ++i;
    return i;
}

// ---- exception path -----------------------------------------------------------------------------
struct ReadRegs {            // a read with <= FC_MAXJ junctions and <= FC_MAXB blocks
    uint32_t jlv[FC_MAXJ], jrv[FC_MAXJ];
    int32_t bsv[FC_MAXB]; uint32_t bev[FC_MAXB];
    uint32_t nj, nb, jrel;
    __device__ __forceinline__ uint32_t jl(uint32_t x) const { return jlv[x]; }
    __device__ __forceinline__ uint32_t jr(uint32_t x) const { return jrv[x]; }
    __device__ __forceinline__ int32_t bs(uint32_t x) const { return bsv[x]; }
    __device__ __forceinline__ uint32_t be(uint32_t x) const { return bev[x]; }
};
template <bool STAGED>
struct ReadWalk {            // any read: every access walks the CIGAR again (long reads only)
    CigSrc<STAGED> cw;
    uint32_t c0, nop;
    int32_t pos;
    uint32_t nj, nb, jrel;
    // x-th operator of the wanted kind (N or mapped): start position, end position, "something advanced before it"
    __device__ __forceinline__ void find(bool want_n, uint32_t x, int32_t& a, int32_t& b, bool& seen) const {
        int32_t cur = pos; uint32_t cnt = 0; bool sn = false;
        a = b = 0; seen = false;
        for (uint32_t q = 0; q < nop; ++q) {
            const uint32_t w = cw(c0 + q), op = w & 15u;
            const int32_t len = (int32_t)(w >> 4);
            const bool isM = op == 0u || op == 7u || op == 8u, isN = op == 3u;
            if ((want_n && isN) || (!want_n && isM)) {
                if (cnt == x) { a = cur; b = cur + len; seen = sn; return; }
                ++cnt;
            }
            if (isM || isN || op == 2u) { cur += len; sn = true; }
        }
    }
    __device__ __forceinline__ uint32_t jl(uint32_t x) const { int32_t a, b; bool s; find(true, x, a, b, s); return (uint32_t)(a - 1) | (s ? 0u : 0x80000000u); }
    __device__ __forceinline__ uint32_t jr(uint32_t x) const { int32_t a, b; bool s; find(true, x, a, b, s); return (uint32_t)(b - 1); }
    __device__ __forceinline__ int32_t bs(uint32_t x) const { int32_t a, b; bool s; find(false, x, a, b, s); return a; }
    __device__ __forceinline__ uint32_t be(uint32_t x) const { int32_t a, b; bool s; find(false, x, a, b, s); return (uint32_t)b; }
};

// one hot (record, operator, side): find the sites the junction is a partner/competitor pair for and classify the read there
template <bool STAGED>
__device__ __noinline__ void hot_item(unsigned long long item, const FStage& st, const FMeta& m, const CigSrc<STAGED>& cw, const FArgs& A) {
    const int anchor = (int)(uint32_t)(item >> 32);
    const uint32_t lo = (uint32_t)item;
    const uint32_t ri = lo >> 21, side = (lo >> 20) & 1u, j = lo & 0xfffffu;
    const int32_t pos = st.pos[m.skip + ri];
    const uint32_t c0 = st.off[m.skip + ri], nop = st.off[m.skip + ri + 1] - c0;
    const uint32_t k = read_class(st.flag[m.skip + ri], A.mode);
    const bool combine = (A.mode & FLAG_COMBINE) != 0;
    ReadRegs rr;
    rr.nj = rr.nb = rr.jrel = 0;
    int32_t cur = pos, jl = 0, jr = 0;
    bool seen = false;
    for (uint32_t q = 0; q < nop; ++q) {
        const uint32_t w = cw(c0 + q), op = w & 15u;
        const int32_t len = (int32_t)(w >> 4);
        if (op == 0u || op == 7u || op == 8u) {
            if (rr.nb < (uint32_t)FC_MAXB) { rr.bsv[rr.nb] = cur; rr.bev[rr.nb] = (uint32_t)(cur + len); }
            ++rr.nb; cur += len; seen = true;
        } else if (op == 3u) {
            if (q == j) { rr.jrel = rr.nj; jl = cur - 1; jr = cur + len - 1; }
            if (rr.nj < (uint32_t)FC_MAXJ) { rr.jlv[rr.nj] = (uint32_t)(cur - 1) | (seen ? 0u : 0x80000000u); rr.jrv[rr.nj] = (uint32_t)(cur + len - 1); }
            ++rr.nj; cur += len; seen = true;
        } else if (op == 2u) {
            cur += len; seen = true;
        }
    }
    const DevGraph& g = A.g;
    if (rr.nj <= (uint32_t)FC_MAXJ && rr.nb <= (uint32_t)FC_MAXB) {
        for (int q = g.rp_off[anchor]; q < g.rp_off[anchor + 1]; ++q) {
            const int t = k4_pair_site(g, q, (int)side, jl, jr);
            if (t >= 0) k4_classify(rr, g, A.cnt, t, k, combine);
        }
    } else {
        const ReadWalk<STAGED> rw{cw, c0, nop, pos, rr.nj, rr.nb, rr.jrel};
        for (int q = g.rp_off[anchor]; q < g.rp_off[anchor + 1]; ++q) {
            const int t = k4_pair_site(g, q, (int)side, jl, jr);
            if (t >= 0) k4_classify(rw, g, A.cnt, t, k, combine);
        }
    }
}

template <bool STAGED>
__device__ __forceinline__ void flush_list(unsigned long long* list, uint32_t& list_n, const FStage& st, const FMeta& m,
                                           const CigSrc<STAGED>& cw, const FArgs& A, int lane) {
    __syncwarp();
    for (uint32_t x = (uint32_t)lane; x < list_n; x += 32) hot_item<STAGED>(list[x], st, m, cw, A);
    __syncwarp();
    list_n = 0;
}

// ---- the records of one staged chunk ---------------------------------------------------------------
template <bool STAGED>
__device__ __forceinline__ void consume(const FStage& st, const FMeta& m, unsigned long long* list, const FArgs& A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DevGraph& g = A.g;
    const CigSrc<STAGED> cw{STAGED ? st.cig : A.rec.cigar, STAGED ? m.cig_base : 0u};
    const int32_t* __restrict__ sp = g.site_pos;
    const uint8_t* __restrict__ hot = g.site_hot;
    const int s0 = m.s0, s1 = m.s1, own_lo = g.own_lo, own_hi = g.own_hi;
    uint32_t list_n = 0;                                             // warp-uniform
    for (uint32_t base = (uint32_t)warp * 32u; base < m.n_rec; base += FC_CONSUMERS) {
        const uint32_t i = base + (uint32_t)lane;
        const bool live = i < m.n_rec;
        int32_t cur = 0;
        uint32_t c0 = 0, nop = 0, k = 0;
        if (live) {
            cur = st.pos[m.skip + i];
            c0 = st.off[m.skip + i];
            nop = st.off[m.skip + i + 1] - c0;
            k = read_class(st.flag[m.skip + i], A.mode);
        }
        const uint32_t maxop = __reduce_max_sync(0xffffffffu, nop);
        int idx = (live && nop) ? bin_lower(g, m, cur, s0) : s0;      // first site with position >= cur, carried along the read
        for (uint32_t j = 0; j < maxop; ++j) {
            if (list_n > (uint32_t)(FC_LIST - 64)) flush_list<STAGED>(list, list_n, st, m, cw, A, lane);
            const uint32_t w = j < nop ? cw(c0 + j) : 5u;             // filler: a zero-length H (no progression)
            const uint32_t op = w & 15u;
            const int32_t len = (int32_t)(w >> 4);
            const bool isM = op == 0u || op == 7u || op == 8u;        // M = X: mapped + advance (S:457-459)
            const bool isN = op == 3u;                                // N: advance, junction (S:480-483)
            const bool adv = isM || isN || op == 2u;                  // D: advance only (S:460-462); I S H P: no progression
            int ie = idx, inx = idx;
            bool v = false;
            uint32_t key_lo = 0, key_hi = 0, hl = 0, hr = 0;
            if (adv && len > 0) {
                const int32_t e = cur + len - 1;                      // last base of the operator
                if (len > FC_WALK) ie = bin_lower(g, m, e, idx);
                while (ie < s1 && __ldg(sp + ie) < e) ++ie;           // first site with position >= e
                inx = ie;
                while (inx < s1 && __ldg(sp + inx) == e) ++inx;       // first site with position > e = lb(cur + len)
                if (isM || isN) {
                    // stabbed positions [cur, e - 1]: covered by the block (a <= t, b >= t + 2, S:469) or strictly inside the
                    // junction (l < t < r with l = cur - 1, r = e; S:507)
                    const int lo = max(idx, own_lo), hi = min(ie, own_hi);
                    v = lo < hi;
                    const uint32_t kind = (isN ? 2u : 0u) + k;
                    key_lo = 4u * (uint32_t)lo + kind; key_hi = 4u * (uint32_t)hi + kind;
                }
            }
            if (isN) {                                                // hot endpoints: l = cur - 1, r = cur + len - 1
                const int32_t l = cur - 1;
                int a = idx;
                while (a > s0 && __ldg(sp + a - 1) == l) --a;         // first site at position l, if any
                if (a < idx && __ldg(hot + a)) hl = (uint32_t)a + 1u;
                if (len > 0) { if (ie < s1 && __ldg(sp + ie) == l + len && __ldg(hot + ie)) hr = (uint32_t)ie + 1u; }
                else hr = hl;                                         // zero-length N: both ends on the same position
            }
            if (__any_sync(0xffffffffu, v)) {
                run_add(A.cnt.diff, key_lo, v, false, lane);
                run_add(A.cnt.diff, key_hi, v, true, lane);
            }
            const uint32_t pm_l = __ballot_sync(0xffffffffu, hl != 0u), pm_r = __ballot_sync(0xffffffffu, hr != 0u);
            if (pm_l | pm_r) {
                const uint32_t lt = (1u << lane) - 1u;
                const uint32_t tag = (i << 21) | (j & 0xfffffu);
                if (hl) list[list_n + __popc(pm_l & lt)] = ((unsigned long long)(hl - 1u) << 32) | tag;
                if (hr) list[list_n + __popc(pm_l) + __popc(pm_r & lt)] = ((unsigned long long)(hr - 1u) << 32) | tag | (1u << 20);
                list_n += (uint32_t)(__popc(pm_l) + __popc(pm_r));
            }
            if (adv) { cur += len; idx = inx; }
        }
    }
    if (list_n) flush_list<STAGED>(list, list_n, st, m, cw, A, lane);
}

__global__ void __launch_bounds__(FC_THREADS, 2) k_count_fused(const __grid_constant__ FArgs A) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FSmem& sm = *reinterpret_cast<FSmem*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < FC_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], FC_CWARPS); }
    }
    __syncthreads();
    if (warp == FC_CWARPS) {
        // ===== producer =====
        if (lane != 0) return;
        uint32_t it = 0;
        const uint32_t n_chunks = A.chunk_hi - A.chunk_lo;
        const uint4* cq = reinterpret_cast<const uint4*>(A.chunks + A.chunk_lo);
        uint32_t nbase = atomicAdd(A.work, (uint32_t)FC_BATCH);
        uint4 nq[FC_BATCH][3];
#pragma unroll
        for (int b = 0; b < FC_BATCH; ++b) {
            const uint32_t c = min(nbase + b, n_chunks - 1u);
#pragma unroll
            for (int q = 0; q < 3; ++q) nq[b][q] = __ldg(cq + 3u * c + q);
        }
        for (;;) {
            const uint32_t base = nbase;
            if (base >= n_chunks) break;
            uint4 cur[FC_BATCH][3];
#pragma unroll
            for (int b = 0; b < FC_BATCH; ++b)
#pragma unroll
                for (int q = 0; q < 3; ++q) cur[b][q] = nq[b][q];
            // the next claim and its descriptors are in flight while this batch is staged
            nbase = atomicAdd(A.work, (uint32_t)FC_BATCH);
#pragma unroll
            for (int b = 0; b < FC_BATCH; ++b) {
                const uint32_t c = min(nbase + b, n_chunks - 1u);
#pragma unroll
                for (int q = 0; q < 3; ++q) nq[b][q] = __ldg(cq + 3u * c + q);
            }
#pragma unroll
            for (int b = 0; b < FC_BATCH; ++b) {
                if (base + b >= n_chunks) continue;
                const uint4 q0 = cur[b][0], q1 = cur[b][1], q2 = cur[b][2];
                const int32_t chrom = (int32_t)q0.x;
                const uint32_t rec_lo = q0.y, rec_hi = q0.z, c_lo = q0.w, c_hi = q1.x;
                const int32_t s0 = (int32_t)q1.y, s1 = (int32_t)q1.z, sb_g0 = (int32_t)q1.w, sb_nb = (int32_t)q2.x;
                if (rec_hi <= rec_lo || s1 <= s0 || c_hi <= c_lo) continue;      // nothing to count on a chromosome without sites
                const uint32_t a0 = rec_lo & ~7u, nr = (rec_hi - a0 + 7u) & ~7u, noff = (rec_hi - a0 + 1u + 3u) & ~3u;
                const uint32_t ca = c_lo & ~3u, nw = (c_hi - ca + 3u) & ~3u;
                const bool staged = nw <= (uint32_t)FC_CIG;
                const uint32_t stage = it % FC_STAGES, parity = (it / FC_STAGES) & 1u;
                mbar_wait_backoff(&sm.empty[stage], parity ^ 1u);
                FMeta& m = sm.meta[stage];
                m.flags = staged ? 0u : FM_GLOBAL_CIG; m.n_rec = rec_hi - rec_lo; m.skip = rec_lo - a0; m.cig_base = ca;
                m.chrom = chrom; m.s0 = s0; m.s1 = s1; m.sb_g0 = sb_g0; m.sb_nb = sb_nb;
                FStage& st = sm.st[stage];
                mbar_expect_tx(&sm.full[stage], nr * 4u + noff * 4u + nr * 2u + (staged ? nw * 4u : 0u));
                bulk_g2s(st.pos, A.rec.pos + a0, nr * 4u, &sm.full[stage]);
                bulk_g2s(st.off, A.rec.cig_off + a0, noff * 4u, &sm.full[stage]);
                bulk_g2s(st.flag, A.rec.flag + a0, nr * 2u, &sm.full[stage]);
                if (staged) bulk_g2s(st.cig, A.rec.cigar + ca, nw * 4u, &sm.full[stage]);
                ++it;
            }
        }
        const uint32_t stage = it % FC_STAGES, parity = (it / FC_STAGES) & 1u;
        mbar_wait_backoff(&sm.empty[stage], parity ^ 1u);
        sm.meta[stage].flags = FM_DONE;
        mbar_arrive(&sm.full[stage]);
        return;
    }
    // ===== consumers =====
    for (uint32_t it = 0;; ++it) {
        const uint32_t stage = it % FC_STAGES, parity = (it / FC_STAGES) & 1u;
        mbar_wait(&sm.full[stage], parity);
        const FMeta m = sm.meta[stage];
        if (m.flags & FM_DONE) break;
        if (m.flags & FM_GLOBAL_CIG) consume<false>(sm.st[stage], m, sm.list[warp], A);
        else consume<true>(sm.st[stage], m, sm.list[warp], A);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[stage]);
    }
}

// CIGAR range and chromosome site / bin ranges of every chunk (runs once the site table is on the device)
__global__ void k_chunk_bounds(FChunk* chunks, uint32_t lo, uint32_t hi, const uint32_t* __restrict__ cig_off, DevGraph g) {
    const uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    FChunk c = chunks[i];
    c.c_lo = cig_off[c.rec_lo]; c.c_hi = cig_off[c.rec_hi];
    if (c.chrom >= 0 && c.chrom < g.n_chrom && g.n_sites > 0) {
        c.s0 = g.cs_off[c.chrom]; c.s1 = g.cs_off[c.chrom + 1];
        c.sb_g0 = g.sb_base[c.chrom]; c.sb_nb = g.sb_base[c.chrom + 1] - c.sb_g0 - 1;
    } else {
        c.s0 = c.s1 = 0; c.sb_g0 = 0; c.sb_nb = 0;
    }
    chunks[i] = c;
}

int fused_grid() {
    static std::mutex mu;
    static int grid_of[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && grid_of[dev]) return grid_of[dev];
    cudaFuncSetAttribute((const void*)k_count_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FSmem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_count_fused, FC_THREADS, sizeof(FSmem));
    if (per_sm < 1) per_sm = 1;
    const int grid = sm_count_current_device() * per_sm;     // one resident CTA per slot: a multiple of the SM count
    if (dev >= 0 && dev < 64) grid_of[dev] = grid;
    return grid;
}

}  // namespace

void launch_chunk_bounds(FChunk* chunks, uint32_t lo, uint32_t hi, const uint32_t* cig_off, DevGraph g, void* stream) {
    if (hi > lo) k_chunk_bounds<<<(hi - lo + 255) / 256, 256, 0, (cudaStream_t)stream>>>(chunks, lo, hi, cig_off, g);
}

// `work` points at a zeroed u32 (the chunk counter of this launch)
void launch_count_fused(const DevRecords& rec, const FChunk* chunks, uint32_t lo, uint32_t hi, DevGraph g, DevCounters cnt,
                        uint32_t* work, uint32_t flags, void* stream) {
    if (hi <= lo || g.n_sites <= 0) return;
    FArgs a{rec, chunks, lo, hi, g, cnt, work, flags};
    const int grid = (int)min((uint32_t)fused_grid(), (hi - lo + (uint32_t)FC_BATCH - 1u) / (uint32_t)FC_BATCH);
    k_count_fused<<<grid, FC_THREADS, sizeof(FSmem), (cudaStream_t)stream>>>(a);
}

}  // namespace spl
