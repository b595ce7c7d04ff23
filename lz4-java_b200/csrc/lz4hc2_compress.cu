// lz4hc2_compress.cu — HC compression, second design (EXPERIMENTAL, selected by b200lz4_hc_algo = 2; the default
// stays lz4hc_compress.cu).  Written at the end of round 1 after the GPU budget was spent: its logic is verified on
// the CPU (tests/simt emulator + the design model tools/study/hc_design_study.c), its speed has not been measured yet.
//
// Same job as lz4hc_compress.cu (LZ4_compress_HC, lz4hc.c:958-973 -> 553-788), re-shaped the way the fast compressor
// was: take the serial parser out of the search.
//
//   K1 search   one block per CTA of 16 warps, 2048 x 32 rings of recent positions in shared memory (the structure of
//               lz4hc_compress.cu).  The block is walked in super-chunks of 512 positions: every thread first inserts
//               its position into its bucket ring, then — after one CTA barrier — searches its own position against
//               the up-to-32 entries of its bucket (entries at or after the position are skipped by the distance
//               arithmetic), per-lane extension up to 64 bytes, longer ones finished by the whole warp.  EVERY position
//               is searched, none waits for a parse decision: pure throughput work, 16 warps in flight per SM.
//               Result: one u32 per position, (length << 16) | distance, length capped at 65535.
//   K2 parse    one THREAD per block.  Backward two-term recurrence over the per-position results,
//                   cost[p] = min(literal: cost[p+1] + 1, match: cost[p + L[p]] + 3 + length bytes),
//               (the recurrence lz4hc.c's level 10+ "optimal" parser approximates; the CPU model shows it beats the
//               lazy rule of level 9 on the same search results), then a forward pass that turns the decisions into
//               sequence records.  Sequential in p, a few operations per position, every block in parallel.
//   K3 layout   one warp per block: sizes, prefix sum of output offsets, 32 tokens/offsets per instruction,
//               lane-parallel literal copies, last literals, result (the layout phase of the fast compressor).
//
// Scratch per block: 4 B (results) + 4 B (costs) per input byte + 16 B per sequence, in a per-thread arena reused
// across sub-batches.  Blocks longer than b200lz4_hc2_max_block bytes are refused (result 0) — the default kernel
// has no such limit.
#include "common.cuh"
#include "kernels.h"
#include "lz4_emit.cuh"
#include <algorithm>

namespace b200 {

static constexpr int HC2_BL = 11, HC2_WAYS = 32, HC2_THREADS = 512, HC2_LANE_CAP = 64, HC2_MAX_ML = 65535;
static constexpr size_t HC2_SMEM = (size_t(2) << HC2_BL) * HC2_WAYS + (size_t(4) << HC2_BL);

__device__ __forceinline__ uint32_t hc2_hash(uint32_t seq) { return (seq * 2654435761u) >> (32 - HC2_BL); }

// ------------------------------------------------------------------------------------------------ K1: search
__global__ void __launch_bounds__(HC2_THREADS, 1)
lz4hc2_search_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                     const int32_t* __restrict__ src_len, uint32_t first, uint32_t nblocks,
                     uint32_t* __restrict__ best_arena, uint32_t stride)
{
    B200_DYN_SMEM(smem_raw, 16);
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem_raw);                              // [bucket][way]
    uint32_t* head = reinterpret_cast<uint32_t*>(smem_raw + (size_t(2) << HC2_BL) * HC2_WAYS);   // [bucket] insertions so far
    if (blockIdx.x >= nblocks) return;
    const uint32_t b = first + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint8_t* __restrict__ src = src_base + src_off[b];
    const int n = src_len[b];
    uint32_t* __restrict__ best = best_arena + size_t(blockIdx.x) * stride;
    if (n < 13 || uint32_t(n) > stride) return;                                          // K2 makes these all-literal / refused
    const InGlobal in{src};
    const int mflimit = n - 12, matchlimit = n - 5;                                      // lz4hc.c:566-567
    for (int i = tid; i < (1 << HC2_BL); i += HC2_THREADS) head[i] = 0;
    __syncthreads();
    for (int c0 = 0; c0 <= mflimit; c0 += HC2_THREADS) {
        const int p = c0 + tid;
        const bool valid = p <= mflimit;
        uint32_t seq = 0, h = 0;
        if (valid) {                                                                     // LZ4HC_Insert (lz4hc.c:120-141), all positions
            seq = in.ld4(p); h = hc2_hash(seq);
            const uint32_t slot = atomicAdd(&head[h], 1u) & (HC2_WAYS - 1);
            ring[h * HC2_WAYS + slot] = uint16_t(p);
        }
        __syncthreads();
        // Two candidates that do not come from the ring.  When more than 32 positions of one super-chunk share a bucket
        // (runs, short periods) the ring ends up holding only the LAST 32 of them — future positions for most of the
        // super-chunk.  In exactly that situation the best source is the nearest earlier position with the same hash,
        // and it is a neighbour: the nearest lane below with an equal hash (one MATCH.ANY), and position p - 1.
        const unsigned same = __match_any_sync(B200_FULL, valid ? h : 0xFFFFFFFFu) & ((1u << lane) - 1u);
        const int near_dist = same ? lane - (31 - __clz((int)same)) : 0;
        int bml = 0, bdist = 0;
        if (valid) {
            const int cnt = (int)min(head[h], (uint32_t)HC2_WAYS);
            const int maxlen = min(matchlimit - p, HC2_LANE_CAP);
            for (int w = -2; w < cnt; w++) {
                const int dist = w == -2 ? near_dist : w == -1 ? 1
                               : int((uint32_t(p) - ring[h * HC2_WAYS + w]) & 0xFFFFu);  // window-relative; aliases are re-verified on the bytes
                const int cand = p - dist;
                if (dist == 0 || cand < 0) continue;                                     // itself, a later position of this super-chunk, or an alias
                if (in.ld4_far(cand) != seq) continue;
                if (bml >= 8 && in.ld4(p + bml - 3) != in.ld4_far(cand + bml - 3)) continue;   // cannot beat the best so far (cf. lz4hc.c:288)
                int ml = 4;
                while (ml < maxlen) {
                    const uint32_t x = in.ld4(p + ml) ^ in.ld4_far(cand + ml);
                    if (x) { ml += (__ffs(x) - 1) >> 3; break; }
                    ml += 4;
                }
                ml = min(ml, maxlen);
                if (ml > bml || (ml == bml && dist < bdist)) { bml = ml; bdist = dist; }
            }
            if (bml < 4) { bml = 0; bdist = 0; }
        }
        // matches that hit the per-lane cap: the whole warp finishes the count, one candidate after the other
        for (unsigned lm = __ballot_sync(B200_FULL, valid && bml >= HC2_LANE_CAP && p + bml < matchlimit); lm; lm &= lm - 1) {
            const int j = __ffs(lm) - 1;
            const int jp = __shfl_sync(B200_FULL, p, j), jml = __shfl_sync(B200_FULL, bml, j), jd = __shfl_sync(B200_FULL, bdist, j);
            const int room = min(matchlimit - (jp + jml), HC2_MAX_ML - jml);
            const int more = match_extend(in, jp + jml, jp - jd + jml, room, lane);
            if (lane == j) bml += more;
        }
        if (valid) best[p] = (uint32_t(bml) << 16) | uint32_t(bdist);
        __syncthreads();                                        // the rings are read until here; the next super-chunk inserts
    }
}

// ------------------------------------------------------------------------------------------------ K2: parse
struct Hc2Rec { int ms, dist, ml, pad; };                       // one sequence: literals up to ms, then the match

__global__ void __launch_bounds__(128)
lz4hc2_parse_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                    const int32_t* __restrict__ src_len, uint32_t first, uint32_t nblocks,
                    uint32_t* __restrict__ best_arena, uint32_t* __restrict__ cost_arena, uint32_t stride,
                    Hc2Rec* __restrict__ rec_arena, uint32_t rec_stride, int32_t* __restrict__ cnt)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    const uint32_t b = first + i;
    const uint8_t* __restrict__ src = src_base + src_off[b];
    const int n = src_len[b];
    if (n < 0 || uint32_t(n) > stride) { cnt[i] = -1; return; }                         // refused: K3 reports 0
    if (n < 13) { cnt[i] = 0; return; }                                                  // all literals (lz4hc.c:566: no match can start)
    uint32_t* __restrict__ best = best_arena + size_t(i) * stride;
    uint32_t* __restrict__ cost = cost_arena + size_t(i) * stride;
    Hc2Rec* __restrict__ rec = rec_arena + size_t(i) * rec_stride;
    const int mflimit = n - 12;
    // costs in 1/256 byte: a literal is 1 byte + its share of the run-length bytes, a match 3 bytes + its length bytes
    constexpr uint32_t LIT = 257, MATCH = 3 * 256;
    uint32_t cnext = uint32_t(n - (mflimit + 1)) * LIT;                                  // the tail behind mflimit is literals
    for (int p = mflimit; p >= 0; p--) {
        const uint32_t e = best[p];
        const int L = int(e >> 16);
        uint32_t c = cnext + LIT;
        bool take = false;
        if (L >= 4) {
            const int q = p + L;
            const uint32_t cq = q > mflimit ? uint32_t(n - q) * LIT : cost[q];
            const uint32_t cm = MATCH + (L - 4 >= 15 ? uint32_t((L - 19) / 255 + 1) * 256u : 0u) + cq;
            if (cm < c) { c = cm; take = true; }
        }
        cost[p] = c; cnext = c;
        if (!take && L) best[p] = 0;
    }
    int k = 0, anchor = 0, p = 0;
    while (p <= mflimit) {
        const uint32_t e = best[p];
        const int L = int(e >> 16);
        if (L == 0) { p++; continue; }
        const int dist = int(e & 0xFFFFu);
        int back = 0;                                                                    // LZ4HC_countBack (lz4hc.c:146-158), at most 8 bytes
        while (back < 8 && p - back > anchor && p - dist - back > 0 && src[p - back - 1] == src[p - dist - back - 1]) back++;
        if (uint32_t(k) < rec_stride) rec[k] = Hc2Rec{p - back, dist, L + back, 0};
        k++;
        p = anchor = p + L;
    }
    cnt[i] = uint32_t(k) <= rec_stride ? k : -1;
}

// ------------------------------------------------------------------------------------------------ K3: layout
__global__ void __launch_bounds__(128)
lz4hc2_layout_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                     const int32_t* __restrict__ src_len,
                     uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                     const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t first, uint32_t nblocks,
                     const Hc2Rec* __restrict__ rec_arena, uint32_t rec_stride, const int32_t* __restrict__ cnt)
{
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nblocks) return;
    const uint32_t b = first + i;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b], cap = dst_cap[b];
    const int total_recs = cnt[i];
    if (total_recs < 0 || n < 0) { if (lane == 0) result[b] = 0; return; }
    if (n == 0) { if (lane == 0) { if (cap >= 1) dst[0] = 0; result[b] = cap >= 1 ? 1 : 0; } return; }
    const Hc2Rec* __restrict__ rec = rec_arena + size_t(i) * rec_stride;
    int op = 0, prev_end = 0; bool fail = false;
    for (int r0 = 0; r0 < total_recs && !fail; r0 += 32) {
        const int cntk = min(32, total_recs - r0);
        Hc2Rec r{0, 1, 4, 0};
        if (lane < cntk) r = rec[r0 + lane];
        const int end = r.ms + r.ml;
        int pe = __shfl_up_sync(B200_FULL, end, 1);
        if (lane == 0) pe = prev_end;
        const int lit = lane < cntk ? r.ms - pe : 0, mcode = r.ml - 4;
        const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
        const int mhdr = mcode >= 15 ? (mcode - 15) / 255 + 1 : 0;
        const long long size = lane < cntk ? 1LL + lhdr + lit + 2 + mhdr : 0LL;
        long long incl = size;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(B200_FULL, incl, d); if (lane >= d) incl += y; }
        const long long total = __shfl_sync(B200_FULL, incl, 31);
        prev_end = __shfl_sync(B200_FULL, end, cntk - 1);
        if ((long long)op + total > (long long)cap) { fail = true; break; }              // lz4hc.c:505-510: output too small
        const int o = op + int(incl - size);
        if (lane < cntk) {
            uint8_t* d = dst + o;
            d[0] = uint8_t((min(lit, 15) << 4) | min(mcode, 15));
            d += 1;
            if (lit >= 15) { int v = lit - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
            d += lit;
            d[0] = uint8_t(r.dist); d[1] = uint8_t(r.dist >> 8);
            d += 2;
            if (mcode >= 15) { int v = mcode - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
        }
        // literals: 16 bytes by the owning lane, longer runs finished by the warp
        const int sn = min(lit, 16);
        if (lane < cntk) for (int t = 0; t < sn; t++) dst[o + 1 + lhdr + t] = src[pe + t];
        for (unsigned lm = __ballot_sync(B200_FULL, lit > 16); lm; lm &= lm - 1) {
            const int j = __ffs(lm) - 1;
            const int ka = __shfl_sync(B200_FULL, pe, j), kl = __shfl_sync(B200_FULL, lit, j), ko = __shfl_sync(B200_FULL, o + 1 + lhdr, j);
            warp_copy(dst + ko + 16, src + ka + 16, kl - 16, lane);
        }
        op += int(total);
    }
    int ret = 0;
    if (!fail) {                                                                         // last literals (lz4hc.c:737-770)
        const int lit = n - prev_end;
        const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
        if ((long long)op + 1 + lhdr + lit <= (long long)cap) {
            if (lane == 0) dst[op] = uint8_t(min(lit, 15) << 4);
            op += 1;
            if (lhdr) { write_len_ext(dst + op, lit - 15, lhdr, lane); op += lhdr; }
            warp_copy(dst + op, src + prev_end, lit, lane);
            ret = op + lit;
        }
    }
    if (lane == 0) result[b] = ret;
}

#ifndef B200_HOST_SIM          // arena and launcher: CUDA only
extern "C" {
int b200lz4_hc2_max_block = 262144;      // longest block the experimental path accepts (sizes its scratch arena)
int b200lz4_hc2_subbatch = 1184;         // blocks per K1/K2/K3 round (148 SMs x 8)
}

struct Hc2Arena { cudaStream_t st; int device; uint32_t blocks, stride; uint32_t* best; uint32_t* cost; Hc2Rec* rec; int32_t* cnt; };

static cudaError_t hc2_arena(cudaStream_t st, uint32_t blocks, uint32_t stride, Hc2Arena** out)
{
    thread_local Hc2Arena ar{nullptr, -1, 0, 0, nullptr, nullptr, nullptr, nullptr};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (ar.device != dev || ar.blocks < blocks || ar.stride < stride) {
        if (ar.best) { cudaFree(ar.best); cudaFree(ar.cost); cudaFree(ar.rec); cudaFree(ar.cnt); ar.best = nullptr; }
        const size_t words = size_t(blocks) * stride, recs = size_t(blocks) * (stride / 4 + 16);
        if ((e = cudaMalloc(&ar.best, words * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ar.cost, words * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ar.rec, recs * sizeof(Hc2Rec))) != cudaSuccess) return e;
        if ((e = cudaMalloc(&ar.cnt, size_t(blocks) * 4)) != cudaSuccess) return e;
        ar.device = dev; ar.blocks = blocks; ar.stride = stride;
    }
    ar.st = st;
    *out = &ar;
    return cudaSuccess;
}

cudaError_t launch_compress_hc2(const BatchArgs& a, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    const uint32_t stride = (uint32_t)((b200lz4_hc2_max_block + 3) & ~3), rec_stride = stride / 4 + 16;
    const uint32_t sub = (uint32_t)std::min<size_t>(a.n, (size_t)std::max(1, b200lz4_hc2_subbatch));
    Hc2Arena* ar = nullptr;
    cudaError_t e = hc2_arena(st, sub, stride, &ar);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(lz4hc2_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HC2_SMEM);
    if (e != cudaSuccess) return e;
    for (size_t first = 0; first < a.n; first += sub) {
        const uint32_t nsb = (uint32_t)std::min<size_t>(sub, a.n - first);
        lz4hc2_search_kernel<<<nsb, HC2_THREADS, HC2_SMEM, st>>>(a.src_base, a.src_off, a.src_len, (uint32_t)first, nsb, ar->best, stride);
        lz4hc2_parse_kernel<<<(nsb + 127) / 128, 128, 0, st>>>(a.src_base, a.src_off, a.src_len, (uint32_t)first, nsb, ar->best, ar->cost, stride,
                                                               ar->rec, rec_stride, ar->cnt);
        lz4hc2_layout_kernel<<<(nsb + 3) / 4, 128, 0, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result,
                                                            (uint32_t)first, nsb, ar->rec, rec_stride, ar->cnt);
    }
    return cudaGetLastError();
}
#endif

} // namespace b200
