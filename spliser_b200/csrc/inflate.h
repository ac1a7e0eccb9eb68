// Raw DEFLATE (RFC 1951) decoder for one BGZF member, written to run as one GPU thread per member (lane 0 of a warp,
// tables in shared memory) and, unchanged, on the host (unit-tested against zlib in tests/test_host.py through
// spl_debug_inflate).  A BGZF member inflates to at most 64 KiB, so the whole LZ77 window is the output buffer itself.
//
// Decoding uses one-level lookup tables of FAST_BITS bits for the literal/length and distance codes; codes longer than
// FAST_BITS fall back to the canonical count/first-code walk (RFC 1951 3.2.2).  Everything is bounds-checked: a corrupt
// member yields an error code, never an out-of-range access.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SPL_HD __host__ __device__ __forceinline__
#else
#define SPL_HD inline
#endif

namespace spl {

constexpr int INF_FAST_BITS = 9;
constexpr int INF_MAXBITS = 15;

// per-decoder scratch (shared memory on the device: one per warp)
struct InflateTables {
    // fast entries: value << 16 | op << 8 | code length; 0 = not a short code.  op: 0 literal / plain symbol (value = symbol),
    // 16 | extra bits = length or distance (value = base), 32 = end of block, 64 = invalid symbol
    uint32_t lit_fast[1 << INF_FAST_BITS];
    uint32_t dist_fast[1 << INF_FAST_BITS];
    uint16_t lit_count[INF_MAXBITS + 1], dist_count[INF_MAXBITS + 1];
    uint16_t lit_sym[288], dist_sym[32];      // symbols ordered by (length, symbol)
    uint8_t  lens[320];                       // code lengths while a dynamic header is read
};

enum : int { INF_OK = 0, INF_ERR_INPUT = 1, INF_ERR_OUTPUT = 2, INF_ERR_CODE = 3, INF_ERR_DIST = 4, INF_ERR_STORED = 5 };

struct BitReader {
    const uint8_t* src;
    uint32_t n, pos;
    uint64_t buf;
    int bits;
    SPL_HD void init(const uint8_t* s, uint32_t len) { src = s; n = len; pos = 0; buf = 0; bits = 0; }
    SPL_HD void refill() {
        if (pos + 4 <= n) {                                 // 4 bytes at once while a whole word of input is left: >= 33 bits
            if (bits > 32) return;                          // afterwards, which covers one literal/length or one distance symbol
            uint32_t w;
#if defined(__CUDA_ARCH__)
            const uint32_t* a = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src + pos) & ~(uintptr_t)3);
            w = __funnelshift_r(a[0], a[1], (uint32_t)(reinterpret_cast<uintptr_t>(src + pos) & 3u) * 8u);   // a[1] stays inside the file image (+ padding)
#else
            w = (uint32_t)src[pos] | ((uint32_t)src[pos + 1] << 8) | ((uint32_t)src[pos + 2] << 16) | ((uint32_t)src[pos + 3] << 24);
#endif
            buf |= (uint64_t)w << bits;
            bits += 32; pos += 4;
            return;
        }
        while (bits <= 56 && pos < n) { buf |= (uint64_t)src[pos++] << bits; bits += 8; }     // tail of the member
    }
    SPL_HD uint32_t peek(int k) const { return (uint32_t)(buf & ((1ull << k) - 1ull)); }
    SPL_HD void drop(int k) { buf >>= k; bits -= k; }
    SPL_HD uint32_t take(int k) { const uint32_t v = peek(k); drop(k); return v; }
};

SPL_HD uint32_t inf_reverse(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

enum : int { INF_KIND_LITLEN = 0, INF_KIND_DIST = 1, INF_KIND_PLAIN = 2 };

// (value << 16 | op << 8) of a symbol: literals and plain symbols carry themselves, length / distance symbols their base
// and extra-bit count in closed form (RFC 1951 3.2.5)
SPL_HD uint32_t inf_entry(int kind, int sym) {
    if (kind == INF_KIND_PLAIN || (kind == INF_KIND_LITLEN && sym < 256)) return (uint32_t)sym << 16;
    if (kind == INF_KIND_LITLEN) {
        if (sym == 256) return 32u << 8;
        if (sym > 285) return 64u << 8;
        const int li = sym - 257;
        const int ext = (li < 8 || li == 28) ? 0 : ((li - 4) >> 2);
        const uint32_t base = li < 8 ? 3u + (uint32_t)li : li == 28 ? 258u : 3u + ((4u + (uint32_t)(li & 3)) << ext);
        return (base << 16) | ((16u | (uint32_t)ext) << 8);
    }
    if (sym > 29) return 64u << 8;
    const int ext = sym < 4 ? 0 : ((sym - 2) >> 1);
    const uint32_t base = sym < 4 ? 1u + (uint32_t)sym : 1u + ((2u + (uint32_t)(sym & 1)) << ext);
    return (base << 16) | ((16u | (uint32_t)ext) << 8);
}

// canonical Huffman tables from code lengths; returns false for an over-subscribed set
SPL_HD bool inf_build(const uint8_t* lens, int n, uint16_t* count, uint16_t* sym, uint32_t* fast, int kind) {
    for (int l = 0; l <= INF_MAXBITS; ++l) count[l] = 0;
    for (int s = 0; s < n; ++s) count[lens[s]]++;
    for (int i = 0; i < (1 << INF_FAST_BITS); ++i) fast[i] = 0;
    int left = 1;
    for (int l = 1; l <= INF_MAXBITS; ++l) {
        left <<= 1;
        left -= (int)count[l];
        if (left < 0) return false;
    }
    uint16_t offs[INF_MAXBITS + 2];
    offs[1] = 0;
    for (int l = 1; l <= INF_MAXBITS; ++l) offs[l + 1] = (uint16_t)(offs[l] + count[l]);
    for (int s = 0; s < n; ++s)
        if (lens[s]) sym[offs[lens[s]]++] = (uint16_t)s;
    // fast table: canonical codes in increasing (length, symbol) order
    uint32_t code = 0;
    int idx = 0;
    for (int l = 1; l <= INF_FAST_BITS; ++l) {
        for (int k = 0; k < (int)count[l]; ++k, ++idx, ++code) {
            const uint32_t rev = inf_reverse(code, l);
            const uint32_t e = inf_entry(kind, sym[idx]) | (uint32_t)l;
            for (uint32_t f = rev; f < (1u << INF_FAST_BITS); f += (1u << l)) fast[f] = e;
        }
        code <<= 1;
    }
    return true;
}

// decode one symbol and consume its code; returns its entry (value << 16 | op << 8), or 64 << 8 on an invalid code
SPL_HD uint32_t inf_decode(BitReader& br, const uint16_t* count, const uint16_t* sym, const uint32_t* fast, int kind) {
    const uint32_t e = fast[br.peek(INF_FAST_BITS)];
    if (e) {
        if ((int)(e & 0xffu) > br.bits) return 64u << 8;
        br.drop((int)(e & 0xffu));
        return e;
    }
    // long code: canonical walk, one bit at a time (RFC 1951 3.2.2 / zlib's puff.c)
    int code = 0, first = 0, index = 0;
    uint64_t b = br.buf;
    for (int len = 1; len <= INF_MAXBITS; ++len) {
        if (len > br.bits) return 64u << 8;
        code |= (int)(b & 1u);
        b >>= 1;
        const int cnt = count[len];
        if (code - cnt < first) { br.drop(len); return inf_entry(kind, sym[index + (code - first)]); }
        index += cnt; first += cnt; first <<= 1; code <<= 1;
    }
    return 64u << 8;
}

// How many cooperating lanes run the decoder.  On the device all 32 lanes of a warp execute the (uniform) decode loop
// redundantly -- same bit buffer, broadcast table reads -- and split the byte copies of LZ77 matches and stored blocks
// between them: a serial copy through global memory pays one load latency PER BYTE, the split copy one per 32 bytes.
// Tables are built by lane 0 alone (the build has read-modify-write steps); everything else the lanes write in common
// is the same value from every lane.
struct OneLane {
    SPL_HD int lane() const { return 0; }
    SPL_HD int count() const { return 1; }
    SPL_HD void sync() const {}
};
#if defined(__CUDACC__)
struct WarpLanes {
    __device__ __forceinline__ int lane() const { return (int)(threadIdx.x & 31u); }
    __device__ __forceinline__ int count() const { return 32; }
    __device__ __forceinline__ void sync() const { __syncwarp(); }
};
#endif

// inflate `src[0..n)` into `dst[0..cap)`; *out_len receives the bytes produced
template <class Lanes>
SPL_HD int inflate_member(const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t cap, uint32_t* out_len, InflateTables& t, const Lanes& ln) {
    const int lane = ln.lane(), nl = ln.count();
    const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    BitReader br;
    br.init(src, n);
    uint32_t out = 0;
    int last = 0;
    while (!last) {
        br.refill();
        if (br.bits < 3) return INF_ERR_INPUT;
        last = (int)br.take(1);
        const uint32_t type = br.take(2);
        if (type == 0) {                                   // stored
            br.drop(br.bits & 7);
            br.refill();
            if (br.bits < 32) return INF_ERR_INPUT;
            const uint32_t len = br.take(16), nlen = br.take(16);
            if ((len ^ 0xffffu) != nlen) return INF_ERR_STORED;
            // bytes still in the bit buffer come first
            uint32_t left = len;
            while (left && br.bits >= 8) {
                if (out >= cap) return INF_ERR_OUTPUT;
                const uint8_t v = (uint8_t)br.take(8);
                if (lane == 0) dst[out] = v;
                ++out; --left;
            }
            if (left) {
                if (br.bits != 0) return INF_ERR_INPUT;
                if (br.pos + left > br.n) return INF_ERR_INPUT;
                if (out + left > cap) return INF_ERR_OUTPUT;
                for (uint32_t i = (uint32_t)lane; i < left; i += (uint32_t)nl) dst[out + i] = br.src[br.pos + i];
                out += left; br.pos += left;
            }
            continue;
        }
        if (type == 3) return INF_ERR_CODE;
        if (type == 1) {                                   // fixed codes
            for (int s = 0; s < 144; ++s) t.lens[s] = 8;
            for (int s = 144; s < 256; ++s) t.lens[s] = 9;
            for (int s = 256; s < 280; ++s) t.lens[s] = 7;
            for (int s = 280; s < 288; ++s) t.lens[s] = 8;
            uint8_t dl[32];
            for (int s = 0; s < 30; ++s) dl[s] = 5;
            ln.sync();
            if (lane == 0) {
                inf_build(t.lens, 288, t.lit_count, t.lit_sym, t.lit_fast, INF_KIND_LITLEN);
                inf_build(dl, 30, t.dist_count, t.dist_sym, t.dist_fast, INF_KIND_DIST);
            }
            ln.sync();
        } else {                                           // dynamic codes
            br.refill();
            if (br.bits < 14) return INF_ERR_INPUT;
            const int nlen = (int)br.take(5) + 257, ndist = (int)br.take(5) + 1, ncode = (int)br.take(4) + 4;
            if (nlen > 286 || ndist > 30) return INF_ERR_CODE;
            uint8_t cl[19];
            for (int i = 0; i < 19; ++i) cl[i] = 0;
            for (int i = 0; i < ncode; ++i) {
                br.refill();
                if (br.bits < 3) return INF_ERR_INPUT;
                cl[ORDER[i]] = (uint8_t)br.take(3);
            }
            // the code-length code reuses the distance tables as scratch
            ln.sync();                                                  // nobody still decodes with the previous block's tables
            if (lane == 0) t.lens[318] = inf_build(cl, 19, t.dist_count, t.dist_sym, t.dist_fast, INF_KIND_PLAIN) ? 1 : 0;   // verdict for every lane:
            ln.sync();                                                  // error exits must be uniform across the warp
            if (!t.lens[318]) return INF_ERR_CODE;
            int idx = 0;
            while (idx < nlen + ndist) {
                br.refill();
                const uint32_t ce = inf_decode(br, t.dist_count, t.dist_sym, t.dist_fast, INF_KIND_PLAIN);
                if (ce & (64u << 8)) return INF_ERR_CODE;
                const int s = (int)(ce >> 16);
                if (s < 16) { t.lens[idx++] = (uint8_t)s; continue; }
                int rep, val = 0;
                if (s == 16) {
                    if (idx == 0) return INF_ERR_CODE;
                    val = t.lens[idx - 1];
                    if (br.bits < 2) return INF_ERR_INPUT;
                    rep = 3 + (int)br.take(2);
                } else if (s == 17) {
                    if (br.bits < 3) return INF_ERR_INPUT;
                    rep = 3 + (int)br.take(3);
                } else {
                    if (br.bits < 7) return INF_ERR_INPUT;
                    rep = 11 + (int)br.take(7);
                }
                if (idx + rep > nlen + ndist) return INF_ERR_CODE;
                while (rep--) t.lens[idx++] = (uint8_t)val;
            }
            if (t.lens[256] == 0) return INF_ERR_CODE;
            // distance lengths first (they sit behind the literal lengths), then the literal/length code
            uint8_t dl[32];
            for (int s = 0; s < ndist; ++s) dl[s] = t.lens[nlen + s];
            ln.sync();
            if (lane == 0) {
                const bool ok = inf_build(t.lens, nlen, t.lit_count, t.lit_sym, t.lit_fast, INF_KIND_LITLEN) &&
                                inf_build(dl, ndist, t.dist_count, t.dist_sym, t.dist_fast, INF_KIND_DIST);
                t.lens[319] = ok ? 1 : 0;                               // slots 318 / 319 are never code lengths (at most 286 + 30)
            }
            ln.sync();
            if (!t.lens[319]) return INF_ERR_CODE;
        }
        // ---- symbols: one table entry gives (literal | length base + extra bits | end of block), a second one the distance
        for (;;) {
            br.refill();
            const uint32_t e = inf_decode(br, t.lit_count, t.lit_sym, t.lit_fast, INF_KIND_LITLEN);
            const uint32_t op = (e >> 8) & 0xffu;
            if (op == 0) {
                if (out >= cap) return INF_ERR_OUTPUT;
                if (lane == 0) dst[out] = (uint8_t)(e >> 16);
                ++out;
                continue;
            }
            if (op & 32u) break;
            if (op & 64u) return INF_ERR_CODE;
            const int lext = (int)(op & 15u);
            if (lext > br.bits) return INF_ERR_INPUT;
            const uint32_t len = (e >> 16) + br.take(lext);
            if (br.bits < 28) br.refill();                     // a distance needs at most 15 + 13 bits
            const uint32_t de = inf_decode(br, t.dist_count, t.dist_sym, t.dist_fast, INF_KIND_DIST);
            if (de & (64u << 8)) return INF_ERR_CODE;
            const int dext = (int)((de >> 8) & 15u);
            if (dext > br.bits) return INF_ERR_INPUT;
            const uint32_t dist = (de >> 16) + br.take(dext);
            if (dist > out) return INF_ERR_DIST;
            if (out + len > cap) return INF_ERR_OUTPUT;
            // every source byte of the match is already written (overlapping matches repeat the last `dist` bytes), so
            // the lanes copy independent bytes; the syncs order them against lane 0's literal stores
            ln.sync();
            if (dist >= len) { for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)nl) dst[out + i] = dst[out - dist + i]; }
            else { for (uint32_t i = (uint32_t)lane; i < len; i += (uint32_t)nl) dst[out + i] = dst[out - dist + (i % dist)]; }
            ln.sync();
            out += len;
        }
    }
    *out_len = out;
    return INF_OK;
}

}  // namespace spl
