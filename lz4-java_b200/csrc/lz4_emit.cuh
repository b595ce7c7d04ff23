// lz4_emit.cuh — pieces shared by the fast and the HC block compressors: input accessors, the
// cooperative match-length counter and the LZ4 sequence writer (token / lengths / literals / offset,
// format: src/lz4/doc/lz4_Block_format.md:25-136 in the reference tree).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace b200 {

// ---- input accessors: the block either stays in global memory (L1/L2-cached) or is staged whole
// into shared memory by a TMA bulk copy (blocks <= 64 KiB); the parser is written once over both.
struct InGlobal {
    const uint8_t* __restrict__ p;
    __device__ __forceinline__ uint32_t ld1(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t ld4(int i) const { return load_u32_unaligned(p + i); }
    // "far" loads = look-backs at candidate positions.  Tried ld.global.cg (bypass L1) for these in
    // round 1: L1 hit rate fell from 44% to 11% and throughput dropped (hl12: 69 -> 59 GiB/s), because the
    // candidate line is re-read by the extension step; they stay on the default cached path.
    __device__ __forceinline__ uint32_t ld1_far(int i) const { return ld1(i); }
    __device__ __forceinline__ uint32_t ld4_far(int i) const { return ld4(i); }
    __device__ __forceinline__ const uint8_t* ptr(int i) const { return p + i; }
};
struct InShared {
    const uint8_t* p;     // generic pointer into the CTA's shared memory
    __device__ __forceinline__ uint32_t ld1(int i) const { return p[i]; }
    __device__ __forceinline__ uint32_t ld4(int i) const {
#ifdef B200_HOST_SIM
        return load_u32_unaligned(p + i);
#endif
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(p) + (uint32_t)i;
        uint32_t lo, hi;
#ifndef B200_HOST_SIM
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(a & ~3u));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"((a & ~3u) + 4u));
#endif
        return __funnelshift_r(lo, hi, (a & 3u) * 8u);
    }
    __device__ __forceinline__ uint32_t ld1_far(int i) const { return ld1(i); }
    __device__ __forceinline__ uint32_t ld4_far(int i) const { return ld4(i); }
    __device__ __forceinline__ const uint8_t* ptr(int i) const { return p + i; }
};

// equal bytes between in[a..] and in[b..] (b < a), at most maxlen; 4 bytes per lane, 128 per round
template <class In>
__device__ __forceinline__ int match_extend(const In& in, int a, int b, int maxlen, int lane)
{
    int total = 0;
    for (;;) {
        const int i = total + lane * 4;
        uint32_t x = 1;                                   // "differs at byte 0" beyond the limit
        if (i < maxlen) x = in.ld4(a + i) ^ in.ld4_far(b + i);
        const unsigned neq = __ballot_sync(B200_FULL, x != 0);
        if (neq) {
            const int fl = __ffs(neq) - 1;
            const uint32_t xf = __shfl_sync(B200_FULL, x, fl);
            const int pos = total + fl * 4 + ((__ffs(xf) - 1) >> 3);
            return min(pos, maxlen);
        }
        total += 128;
    }
}

// 255-chain for a length field whose token nibble saturated: v = length - 15 >= 0, cnt = v/255 + 1 bytes
__device__ __forceinline__ void write_len_ext(uint8_t* d, int v, int cnt, int lane)
{
    for (int i = lane; i < cnt; i += 32) d[i] = (i == cnt - 1) ? uint8_t(v - 255 * (cnt - 1)) : uint8_t(255);
}

// One LZ4 sequence waiting to be written: literals [anchor, ms) then a match of ml bytes at distance off.
struct Seq { int anchor, ms, off, ml; };

// token, [literal length], literals, offset, [match length] with lane-parallel stores.
// Returns false if dst is too small (lz4.c:1085-1088, 1158).
template <class In>
__device__ __forceinline__ bool emit_sequence(const In& in, const Seq& q, uint32_t litv, uint8_t* __restrict__ dst, int& op, int cap, int lane)
{
    const int lit = q.ms - q.anchor;
    const int mcode = q.ml - 4;
    const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
    const int mhdr = mcode >= 15 ? (mcode - 15) / 255 + 1 : 0;
    if ((long long)op + 1 + lhdr + lit + 2 + mhdr > cap) return false;
    uint8_t* d = dst + op;
    if (lane == 0) d[0] = uint8_t((min(lit, 15) << 4) | min(mcode, 15));
    d += 1;
    if (lhdr) { write_len_ext(d, lit - 15, lhdr, lane); d += lhdr; }
    if (lit <= 32) { if (lane < lit) d[lane] = uint8_t(litv); }                // byte preloaded by the caller
    else warp_copy_words<false>(d, in.ptr(q.anchor), lit, lane);
    d += lit;
    if (lane < 2) d[lane] = uint8_t(q.off >> (8 * lane));                       // LE16 offset (lz4.c:1133)
    d += 2;
    if (mhdr) { write_len_ext(d, mcode - 15, mhdr, lane); }
    op += 1 + lhdr + lit + 2 + mhdr;
    return true;
}

} // namespace b200
