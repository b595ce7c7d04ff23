// lz4_compress.cu — batch LZ4 fast block compression, one independent block per warp.
//
// Replaces the reference's LZ4_compress_default (lz4.c:1435 -> 1416 -> 1346 -> 910-1302) as
// called from the JNI shim (src/jni/net_jpountz_lz4_LZ4JNI.c:75).  Same algorithm family — greedy
// single-probe LZ77 over a 4-byte multiplicative hash (lz4.c:756-762), 16-bit block-relative
// position table for blocks < 64 KiB (lz4.c:1353), MFLIMIT/LASTLITERALS end rules (lz4.c:243-244)
// — but re-shaped for a 32-lane warp, so the emitted stream is a *different valid parse* of the
// same format (like the reference's own Java ports, README.md:45-47): it round-trips bit-exactly
// through every LZ4 decoder; its ratio is reported next to the reference's.
//
// Warp algorithm (one "step" = 32 consecutive positions):
//   1. lane l reads the 4 bytes at ip+l, hashes them, fetches the candidate position from the
//      warp's private hash table in shared memory;
//   2. candidates are verified against the block (4-byte compare); __ballot_sync + __ffs picks the
//      FIRST matching position (greedy, like the scalar parse);
//   3. lanes at or before the match start publish their positions to the table (positions after
//      the match start are not published: they will be probed again after the match);
//   4. the match is extended backwards (catch-up, lz4.c:1080) and forwards (LZ4_count,
//      lz4.c:659-682) with lane-parallel compares + ballot;
//   5. token / literal run / offset / length bytes are emitted with lane-parallel stores.
// Without a hit the step costs one pass and advances 32 positions.
//
// Algorithmic HBM bytes per block: N (input read once) + C (output written once).  The table
// (2^HASH_LOG entries) lives in shared memory and never touches HBM.
#include "common.cuh"
#include "kernels.h"
#include "lz4_emit.cuh"
#include "lz4_compress_wide.cuh"
#include <type_traits>
#include <algorithm>

namespace b200 {

// The greedy warp parser.  Returns the compressed size, 0 if dst is too small.
//
// Software-pipelined by one sequence: the probe loads of step k+1 (input window, table, candidate
// bytes) are issued BEFORE sequence k is written out, so the emission stores and literal copies fill
// the latency of those dependent loads instead of adding to the serial chain.
template <int HASH_LOG, bool U16, class In, class Entry>
__device__ __forceinline__ int compress_block(const In in, const uint8_t* __restrict__ gsrc, int n,
                                              uint8_t* __restrict__ dst, int cap, Entry* table, int lane)
{
    int op = 0, anchor = 0, ip = 0;
    const int mflimit = n - 12;        // last position a match may start at (MFLIMIT, lz4.c:243)
    const int matchlimit = n - 5;      // matches end here at the latest (LASTLITERALS, lz4.c:244)
    int pf = 0;                        // software prefetch cursor (global input only)
    bool have = false;                 // a found-but-not-yet-written sequence
    Seq q = {0, 0, 0, 0};

    while (ip <= mflimit) {            // n < 13 never finds a match: all literals (lz4.c:981)
        if (std::is_same<In, InGlobal>::value) {
            if (pf < ip + 2048) {      // keep ~4 KiB of the forward stream on its way to L2
                const int qq = pf + lane * 128;
                if (qq < n) B200_PREFETCH_L2(gsrc + qq);
                pf += 4096;
            }
        }
        // ---- A: probe 32 positions; the pending sequence's literal bytes are fetched alongside
        uint32_t litv = 0;
        if (have && lane < q.ms - q.anchor && q.ms - q.anchor <= 32) litv = in.ld1(q.anchor + lane);
        if (std::is_same<In, InGlobal>::value && lane == 0 && ip + 160 < n)     // next line of the forward stream -> L1
            B200_PREFETCH_L1(gsrc + ip + 128);
        const int p = ip + lane;
        const bool valid = p <= mflimit;
        const int pp = min(p, mflimit);
        const uint32_t seq = in.ld4(pp);
        const uint32_t h = (seq * 2654435761u) >> (32 - HASH_LOG);
        const int cand = table[h];
        const bool plausible = valid && cand < p && (U16 || p - cand <= 65535);
        const uint32_t cseq = in.ld4_far(plausible ? cand : pp);

        // ---- B: write the previous sequence while those loads are in flight
        if (have) {
            if (!emit_sequence(in, q, litv, dst, op, cap, lane)) return 0;
            have = false;
        }

        // ---- C: vote, publish, extend
        const bool hit = plausible && cseq == seq;
        const unsigned m = __ballot_sync(B200_FULL, hit);
        const int f = m ? __ffs(m) - 1 : 31;
        if (valid && lane <= f) table[h] = Entry(p);
        if (m == 0) { ip += 32; continue; }

        int ms = ip + f;                                        // match start
        int mc = __shfl_sync(B200_FULL, cand, f);               // where the same bytes occurred before
        int ml;
        {   // one cooperative round for both directions: lane j compares offset d = j-8, i.e. up to
            // 8 bytes of catch-up behind the match (lz4.c:1080) and its first 24 bytes (lz4.c:1153)
            const int d = lane - 8;
            const int backroom = min(ms - anchor, mc);
            const bool ok = d < 0 ? (-d <= backroom) : (ms + d < matchlimit);
            const bool eq = ok && in.ld1(ms + d) == in.ld1_far(mc + d);
            const unsigned e = __ballot_sync(B200_FULL, eq);
            const int back = __clz((~e) & 0xFFu) - 24;                       // ones below bit 8, contiguous from bit 7
            const int fwd = __ffs((~(e >> 8)) | (1u << 24)) - 1;             // ones from bit 8 upwards, <= 24
            ml = fwd;
            if (fwd == 24) ml += match_extend(in, ms + 24, mc + 24, matchlimit - (ms + 24), lane);
            ms -= back; mc -= back; ml += back;
        }
        q.anchor = anchor; q.ms = ms; q.off = ms - mc; q.ml = ml; have = true;
        ip = anchor = ms + ml;
    }
    if (have) {
        uint32_t litv = 0;
        if (lane < q.ms - q.anchor && q.ms - q.anchor <= 32) litv = in.ld1(q.anchor + lane);
        if (!emit_sequence(in, q, litv, dst, op, cap, lane)) return 0;
    }

    {   // last literals (lz4.c:1266-1293)
        const int lit = n - anchor;
        const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
        if ((long long)op + 1 + lhdr + lit > cap) return 0;
        if (lane == 0) dst[op] = uint8_t(min(lit, 15) << 4);
        op += 1;
        if (lhdr) { write_len_ext(dst + op, lit - 15, lhdr, lane); op += lhdr; }
        warp_copy(dst + op, in.ptr(anchor), lit, lane);
        op += lit;
    }
    return op;
}

static constexpr int STAGE_BYTES = 65536 + 16;     // staged input capacity (U16 blocks are < 65547 bytes)

template <int HASH_LOG, bool U16, bool STAGE>
__global__ void __launch_bounds__(32)
lz4_compress_fast_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                         const int32_t* __restrict__ src_len,
                         uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                         const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t nblocks)
{
    using Entry = typename std::conditional<U16, uint16_t, uint32_t>::type;
    B200_DYN_SMEM(smem_raw, 128);
    Entry* table = reinterpret_cast<Entry*>(smem_raw);
    constexpr int TABLE_BYTES = int(sizeof(Entry) << HASH_LOG);

    const uint32_t b = blockIdx.x;
    if (b >= nblocks) return;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b];
    const int cap = dst_cap[b];
    int ret = 0;

    if (n < 0 || n > 0x7E000000) goto done;                       // lz4.c:1324
    if (U16 && n >= 65536 + 11) goto done;                         // lz4.c:973 (caller broke the max_src_len promise)
    if (n == 0) {                                                  // lz4.c:1325-1336
        if (cap >= 1) { if (lane == 0) dst[0] = 0; ret = 1; }
        goto done;
    }
    for (int i = lane; i < TABLE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
    if (STAGE) {
        // whole block -> shared memory: one TMA bulk copy (16-byte aligned part) + a lane-copied tail
        uint8_t* stage = smem_raw + TABLE_BYTES + 16;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem_raw + TABLE_BYTES);
#ifdef B200_HOST_SIM
        const bool aligned = false;                          // no TMA in the emulator: the lane copy below stages everything
#else
        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
#endif
        const int bulk = aligned ? (n & ~15) : 0;
#ifndef B200_HOST_SIM
        if (lane == 0 && bulk) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bulk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(src), "r"(bulk), "r"(bar) : "memory");
        }
#endif
        for (int i = bulk + lane; i < n; i += 32) stage[i] = src[i];
        __syncwarp();
#ifndef B200_HOST_SIM
        if (bulk) {
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar) : "memory");
            } while (!ok);
        }
#endif
        __syncwarp();
        ret = compress_block<HASH_LOG, U16>(InShared{stage}, src, n, dst, cap, table, lane);
    } else {
        __syncwarp();
        ret = compress_block<HASH_LOG, U16>(InGlobal{src}, src, n, dst, cap, table, lane);
    }
done:
    if (lane == 0) result[b] = ret;
}

// ---------------------------------------------------------------------------------------------
// Decoupled parser (algo 2, the default): candidate lookup + verification run AHEAD of the greedy parse.
//
// The coupled kernel above discovers one match, extends it, writes it, and only then knows where to
// probe next: three dependent L2 round trips and ~200 warp instructions per ~27 input bytes, with 13
// warps/SM to hide them.  Here the block is walked in chunks of 128 positions, each in two phases:
//
//   AB (lane-parallel, no decisions): lane l owns the 4 consecutive positions of one aligned 32-bit
//      word of the input (two coalesced word loads + three funnel shifts give its four 4-byte
//      sequences).  All 128 positions are hashed, looked up and then INSERTED (every position, in
//      order), all candidates are verified with four independent loads per lane, and the outcome is
//      one 16-bit distance per position in shared memory (256 B per warp) plus a 128-bit hit mask in
//      registers.  Nothing here depends on the parse.
//   C  (serial, cheap): the greedy walk reads only that: next hit at or after ip from the mask
//      (a few uniform ALU ops), its distance by one shared-memory read, ONE cooperative compare round
//      for catch-up (lz4.c:1080) + the first 24 match bytes (lz4.c:1153), literal copy, and a record
//      of the sequence in lane k's registers.  Tokens and offsets of 32 recorded sequences are then
//      written by 32 lanes at once.
//
// Inserting every position (instead of only positions outside matches) costs no ratio: on the
// reference's own generator the parse is slightly denser than lz4's (1.63 vs 1.61 at P=0.50).
template <int HASH_LOG, bool U16>
__global__ void __launch_bounds__(32)
lz4_compress_fast2_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                          const int32_t* __restrict__ src_len,
                          uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                          const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t nblocks)
{
    using Entry = typename std::conditional<U16, uint16_t, uint32_t>::type;
    constexpr int TABLE_BYTES = int(sizeof(Entry) << HASH_LOG);
    B200_DYN_SMEM(smem_raw, 128);
    Entry* table = reinterpret_cast<Entry*>(smem_raw);
    uint16_t* s_dist = reinterpret_cast<uint16_t*>(smem_raw + TABLE_BYTES);        // [128] match distance, 0 = no match at this position

    const uint32_t b = blockIdx.x;
    if (b >= nblocks) return;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b];
    const int cap = dst_cap[b];
    int ret = 0;

    if (n < 0 || n > 0x7E000000 || cap < 0) goto done;            // lz4.c:1324; no room at all
    if (U16 && n >= 65536 + 11) goto done;                         // lz4.c:973
    if (n == 0) { if (cap >= 1) { if (lane == 0) dst[0] = 0; ret = 1; } goto done; }
    {
        for (int i = lane; i < TABLE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        // aligned-word view of the block: byte a of the view is position a - ph
        const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(src)) & 3u;
        const uint32_t* __restrict__ wsrc = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) - ph);
        auto ld4 = [&](int pos) -> uint32_t {                      // the 4 bytes at position pos (pos + 3 < n)
            const uint32_t a = uint32_t(pos) + ph;
            const uint32_t* w = wsrc + (a >> 2);
            return __funnelshift_r(w[0], w[1], (a & 3u) * 8u);
        };
        const int mflimit = n - 12, matchlimit = n - 5;            // lz4.c:243-244
        int op = 0, anchor = 0, ip = 0;
        // up to 32 found sequences wait in registers (lane k holds sequence k) for their token/offset bytes
        int nrec = 0, r_o = 0, r_lit = 0, r_ml = 0, r_dist = 0;

        auto flush = [&]() {                                       // 32 lanes write 32 tokens / length chains / offsets
            if (lane < nrec) {
                const int mcode = r_ml - 4;
                uint8_t* d = dst + r_o;
                d[0] = uint8_t((min(r_lit, 15) << 4) | min(mcode, 15));
                d += 1;
                if (r_lit >= 15) { int v = r_lit - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
                d += r_lit;
                d[0] = uint8_t(r_dist); d[1] = uint8_t(r_dist >> 8);                  // LE16 offset (lz4.c:1133)
                d += 2;
                if (mcode >= 15) { int v = mcode - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
            }
            nrec = 0;
        };

        const int nchunks = (mflimit + int(ph)) / 128 + 1;
        for (int c = 0; c < nchunks; c++) {
            const int cp0 = 128 * c - int(ph);                     // position of the chunk's first byte
            if (lane < 2) {                                        // two chunks ahead -> L2
                const int pfq = cp0 + 512 + lane * 128;
                if (pfq < n) B200_PREFETCH_L2(src + pfq);
            }
            // ---------------- phase AB
            const int p0 = cp0 + 4 * lane;
            uint32_t w0 = 0, w1 = 0;
            if (p0 + 3 >= 0 && p0 <= mflimit) { w0 = wsrc[32 * c + lane]; w1 = wsrc[32 * c + lane + 1]; }
            uint32_t seq[4], h[4]; int cand[4]; bool plaus[4];
            seq[0] = w0; seq[1] = __funnelshift_r(w0, w1, 8); seq[2] = __funnelshift_r(w0, w1, 16); seq[3] = __funnelshift_r(w0, w1, 24);
            #pragma unroll
            for (int j = 0; j < 4; j++) { h[j] = (seq[j] * 2654435761u) >> (32 - HASH_LOG); cand[j] = table[h[j]]; }
            __syncwarp();      // every lookup of the chunk precedes every insert.  Lanes whose sequences hash alike then
                               // store to the same slot: any winner is a valid (earlier) position, so the race is benign
            #pragma unroll
            for (int j = 0; j < 4; j++) {
                const int p = p0 + j;
                const bool valid = p >= 0 && p <= mflimit;
                if (valid) table[h[j]] = Entry(p);
                plaus[j] = valid && cand[j] < p && (U16 || p - cand[j] <= 65535);
            }
            uint32_t cseq[4];
            #pragma unroll
            for (int j = 0; j < 4; j++) cseq[j] = plaus[j] ? ld4(cand[j]) : ~seq[j];
            uint32_t nib = 0; uint32_t d01, d23;
            {
                uint32_t dd[4];
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool hit = cseq[j] == seq[j];            // implies plaus (otherwise cseq = ~seq)
                    dd[j] = hit ? uint32_t(p0 + j - cand[j]) : 0u;
                    nib |= uint32_t(hit) << j;
                }
                d01 = dd[0] | (dd[1] << 16); d23 = dd[2] | (dd[3] << 16);
            }
            reinterpret_cast<uint2*>(s_dist)[lane] = make_uint2(d01, d23);
            // position-ordered hit mask: word k = positions cp0+32k .. +31 (8 lanes x 4 bits)
            uint32_t gw = nib << (4 * (lane & 7));            // OR over each group of 8 lanes (butterfly: all lanes in step)
            gw |= __shfl_xor_sync(B200_FULL, gw, 1); gw |= __shfl_xor_sync(B200_FULL, gw, 2); gw |= __shfl_xor_sync(B200_FULL, gw, 4);
            const uint32_t hw0 = __shfl_sync(B200_FULL, gw, 0), hw1 = __shfl_sync(B200_FULL, gw, 8),
                           hw2 = __shfl_sync(B200_FULL, gw, 16), hw3 = __shfl_sync(B200_FULL, gw, 24);
            const unsigned long long hlo = (unsigned long long)hw0 | ((unsigned long long)hw1 << 32);
            const unsigned long long hhi = (unsigned long long)hw2 | ((unsigned long long)hw3 << 32);
            __syncwarp();
            // ---------------- phase C: greedy walk over this chunk's hits
            for (;;) {
                int r = ip - cp0;
                if (r < 0) r = 0;
                if (r >= 128) break;
                int q;
                {   // first hit at or after r: two 64-bit halves of the 128-bit position-ordered mask
                    const unsigned long long lo = r < 64 ? (hlo >> r) : 0ull;
                    if (lo) q = r + __ffsll((long long)lo) - 1;
                    else {
                        const int r2 = max(r - 64, 0);
                        const unsigned long long hi = hhi >> r2;
                        if (hi == 0) break;
                        q = 64 + r2 + __ffsll((long long)hi) - 1;
                    }
                }
                int ms = cp0 + q;
                const int dist = s_dist[q];
                int mc = ms - dist, ml;
                {   // one cooperative round: lane j compares offset d = j-8 (catch-up) .. +23 (match body)
                    const int d = lane - 8;
                    const int backroom = min(ms - anchor, mc);
                    const bool ok = d < 0 ? (-d <= backroom) : (ms + d < matchlimit);
                    const bool eq = ok && src[ms + d] == src[mc + d];
                    const unsigned e = __ballot_sync(B200_FULL, eq);
                    const int back = __clz((~e) & 0xFFu) - 24;
                    const int fwd = __ffs((~(e >> 8)) | (1u << 24)) - 1;
                    ml = fwd;
                    if (fwd == 24) ml += match_extend(InGlobal{src}, ms + 24, mc + 24, matchlimit - (ms + 24), lane);
                    ms -= back; ml += back;
                }
                // sizes, literal copy now (offsets are known sequentially), header bytes later in batch
                const int lit = ms - anchor, mcode = ml - 4;
                const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
                const int mhdr = mcode >= 15 ? (mcode - 15) / 255 + 1 : 0;
                const int size = 1 + lhdr + lit + 2 + mhdr;
                if (uint32_t(op) + uint32_t(size) > uint32_t(cap)) goto done;         // lz4.c:1085-1088, 1158 (op <= cap < 2^31)
                warp_copy(dst + op + 1 + lhdr, src + anchor, lit, lane);
                if (lane == nrec) { r_o = op; r_lit = lit; r_ml = ml; r_dist = dist; }
                nrec++;
                if (nrec == 32) flush();
                op += size;
                ip = anchor = ms + ml;
            }
        }
        flush();
        {   // last literals (lz4.c:1266-1293)
            const int lit = n - anchor;
            const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
            if ((long long)op + 1 + lhdr + lit > cap) goto done;
            if (lane == 0) dst[op] = uint8_t(min(lit, 15) << 4);
            op += 1;
            if (lhdr) { write_len_ext(dst + op, lit - 15, lhdr, lane); op += lhdr; }
            warp_copy(dst + op, src + anchor, lit, lane);
            op += lit;
        }
        ret = op;
    }
done:
    if (lane == 0) result[b] = ret;
}

#ifndef B200_HOST_SIM
template <int HASH_LOG, bool U16>
static cudaError_t launch_v2(const BatchArgs& a, cudaStream_t st)
{
    const size_t smem = ((U16 ? 2u : 4u) << HASH_LOG) + 128 * sizeof(uint16_t);
    auto k = lz4_compress_fast2_kernel<HASH_LOG, U16>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<(unsigned)a.n, 32, smem, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap,
                                       a.result, (uint32_t)a.n);
    return cudaGetLastError();
}
#endif

// ---------------------------------------------------------------------------------------------
// Warp-specialised pipeline (algo 3): the decoupled parser above, split across producer/consumer warps of
// one CTA so the serial walk carries nothing but the walk.
//
//   warp L ("lookup"):  phase AB for chunk c — hash, table probe + insert, candidate verification — and
//                       publishes the chunk's distances (u16 per position) and hit masks in shared memory.
//   warp P ("parser"):  greedy walk of chunk c.  Every lane takes one of the chunk's next 32 hits and measures
//                       it on its own (catch-up of up to 4 bytes, body up to 32 bytes), so 32 extension loads
//                       are in flight at once; each lane then finds its successor (first hit starting at or
//                       after its own end) by a shuffle binary search over the sorted hit positions, and the
//                       greedy chain is one shuffle per selected sequence.  Selected lanes write 8-byte records.
//   warp E ("emit"):    sizes, prefix sum of output offsets, 32 tokens/offsets by 32 lanes, lane-parallel
//                       literal copies, last literals, result.  With B200_V3_WARPS == 2 warp L does this
//                       after its lookup phase (one chunk behind the parser).
// Chunk state is double-buffered; the hand-offs are named barriers (producer bar.arrive, consumer bar.sync),
// so no warp waits for a warp it does not depend on.  Same parse, same output bytes as algo 2.
#ifndef B200_V3_NB
#define B200_V3_NB 2
#endif
#ifndef B200_V3_WARPS
#define B200_V3_WARPS 2
#endif
#ifndef B200_V3_MINB
#define B200_V3_MINB 12
#endif
#ifndef B200_V3_FENCE
#define B200_V3_FENCE 0
#endif
#ifndef B200_V3_MINB12
#define B200_V3_MINB12 16
#endif
// B200_V3_RUNS = 1 (experimental, unmeasured; DESIGN.md "Round-2 plan"): warp L also publishes which hits START a run
// (a hit continues a run when the position before it hit with the same distance), and warp P ranks and measures run
// starts only.  A sequence that has to start inside a run (at the end of the previous one) takes the run's distance and
// ends where the run start's match ends, so nothing is measured twice.  Same greedy parse, same output bytes as 0.
#ifndef B200_V3_RUNS
#define B200_V3_RUNS 0
#endif
// B200_V3_SPLIT = 1 (experimental, unmeasured; needs B200_V3_RUNS and the two-warp build): the parser warp also lays out
// the sequence headers (sizes, prefix sum of output offsets, token / length bytes / offset) while it still holds the
// selected sequences in registers, and hands warp L only (literal source, literal count, destination) triples, so L is
// left with lookup + literal copies.  Same output bytes.
#ifndef B200_V3_SPLIT
#define B200_V3_SPLIT 0
#endif
#if B200_V3_SPLIT && !(B200_V3_RUNS && B200_V3_WARPS == 2)
#error "B200_V3_SPLIT needs B200_V3_RUNS=1 and B200_V3_WARPS=2"
#endif
#ifdef B200_HOST_SIM
__device__ __forceinline__ void bar_arrive(int id) { simt::bar_arrive(id, 64); }
__device__ __forceinline__ void bar_wait(int id) { simt::bar_sync(id, 64); }
#else
// (Immediate barrier ids, so ptxas reserves only the barriers in use and not all 16.)
#define B200_BAR_CASE(OP, N) case N: asm volatile(OP " " #N ", 64;" ::: "memory"); break;
__device__ __forceinline__ void bar_arrive(int id)
{
#if B200_V3_FENCE
    __threadfence_block();                               // order this warp's shared-memory writes before the arrival
#endif
    switch (id) { B200_BAR_CASE("bar.arrive", 1) B200_BAR_CASE("bar.arrive", 2) B200_BAR_CASE("bar.arrive", 3)
                  B200_BAR_CASE("bar.arrive", 4) B200_BAR_CASE("bar.arrive", 5) B200_BAR_CASE("bar.arrive", 6)
                  B200_BAR_CASE("bar.arrive", 7) default: asm volatile("bar.arrive 8, 64;" ::: "memory"); }
}
__device__ __forceinline__ void bar_wait(int id)
{
    switch (id) { B200_BAR_CASE("bar.sync", 1) B200_BAR_CASE("bar.sync", 2) B200_BAR_CASE("bar.sync", 3)
                  B200_BAR_CASE("bar.sync", 4) B200_BAR_CASE("bar.sync", 5) B200_BAR_CASE("bar.sync", 6)
                  B200_BAR_CASE("bar.sync", 7) default: asm volatile("bar.sync 8, 64;" ::: "memory"); }
}

#endif

template <int HASH_LOG, bool SPARSE>
__global__ void __launch_bounds__(32 * B200_V3_WARPS, HASH_LOG == 13 ? B200_V3_MINB : B200_V3_MINB12)
lz4_compress_fast3_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                          const int32_t* __restrict__ src_len,
                          uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                          const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t nblocks)
{
    constexpr int NB = B200_V3_NB, NW = B200_V3_WARPS;
    constexpr int LAG = NB - 1;                            // NW == 2: warp L emits chunk i-LAG after looking chunk i up
    constexpr int BAR_FULL = 1, BAR_WALKED = 1 + NB, BAR_DFREE = 1 + 2 * NB, BAR_RFREE = 1 + 3 * NB;
    static_assert(NW == 2 || 4 * NB <= 8, "named barrier ids 1..8");
    constexpr int TABLE_BYTES = 2 << HASH_LOG;
    B200_DYN_SMEM(smem_raw, 128);
    uint16_t* table = reinterpret_cast<uint16_t*>(smem_raw);
    uint16_t* s_dist = reinterpret_cast<uint16_t*>(smem_raw + TABLE_BYTES);                 // [NB][128] distance per position, 0 = no match
    constexpr int REC_BYTES = B200_V3_SPLIT ? 512 : 256;
#if B200_V3_SPLIT
    uint4* s_rec4 = reinterpret_cast<uint4*>(smem_raw + TABLE_BYTES + NB * 256);            // [NB][32]  x = literal source, y = literal count, z = destination
#else
    uint2* s_rec = reinterpret_cast<uint2*>(smem_raw + TABLE_BYTES + NB * 256);             // [NB][32]  x = start | distance << 16, y = length
#endif
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem_raw + TABLE_BYTES + NB * 256 + NB * REC_BYTES);   // [NB][4] hit masks
    int* s_cnt = reinterpret_cast<int*>(s_mask + 4 * NB);                                   // [NB] records per buffer
    uint8_t* s_hit = reinterpret_cast<uint8_t*>(s_cnt + 4);                                 // [128] ranked hit positions (warp P's scratch)
#if B200_V3_RUNS
    uint32_t* s_rmask = reinterpret_cast<uint32_t*>(s_hit + 128);                           // [NB][4] run-start masks
#endif
#if B200_V3_SPLIT
    int* s_state = reinterpret_cast<int*>(s_rmask + 4 * NB);                                // [NB][4] output offset, end of the last sequence, failed
#endif

    const uint32_t b = blockIdx.x;
    if (b >= nblocks) return;
    const int lane = lane_id();
    const int role = threadIdx.x >> 5;                     // 0 = L, 1 = P, 2 = E
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b];
    const int cap = dst_cap[b];

    if (n < 0 || n >= 65536 + 11) { if (threadIdx.x == 0) result[b] = 0; return; }               // lz4.c:1324, 973
    if (n == 0) { if (threadIdx.x == 0) { if (cap >= 1) dst[0] = 0; result[b] = cap >= 1 ? 1 : 0; } return; }

    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(src)) & 3u;
    const uint32_t* __restrict__ wsrc = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) - ph);
    const int mflimit = n - 12, matchlimit = n - 5;
    const int nchunks = (mflimit + int(ph)) / 128 + 1;
    auto ld4 = [&](int pos) -> uint32_t {
        const uint32_t a = uint32_t(pos) + ph;
        const uint32_t* w = wsrc + (a >> 2);
        return __funnelshift_r(w[0], w[1], (a & 3u) * 8u);
    };

    // ------------------------------------------------------------------ phase AB for chunk c (warp L)
    auto lookup = [&](int c) {
        const int cp0 = 128 * c - int(ph), buf = c % NB;
        if (lane < 2) {
            const int pfq = cp0 + 512 + lane * 128;
            if (pfq < n) B200_PREFETCH_L2(src + pfq);
        }
        const int p0 = cp0 + 4 * lane;
        uint32_t w0 = 0, w1 = 0;
        if (p0 + 3 >= 0 && p0 <= mflimit) { w0 = wsrc[32 * c + lane]; w1 = wsrc[32 * c + lane + 1]; }
        uint32_t seq[4], h[4]; int cand[4]; bool plaus[4];
        seq[0] = w0; seq[1] = __funnelshift_r(w0, w1, 8); seq[2] = __funnelshift_r(w0, w1, 16); seq[3] = __funnelshift_r(w0, w1, 24);
        #pragma unroll
        for (int j = 0; j < 4; j++) { h[j] = (seq[j] * 2654435761u) >> (32 - HASH_LOG); cand[j] = table[h[j]]; }
        __syncwarp();  // every lookup of the chunk precedes every insert (same-slot stores: any winner is a valid position)
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + j;
            const bool valid = p >= 0 && p <= mflimit;
            // SPARSE: only one position in four is published (every position is still probed)
            if (valid && (!SPARSE || j == 0)) table[h[j]] = uint16_t(p);
            plaus[j] = valid && cand[j] < p;
        }
        uint32_t cseq[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) cseq[j] = plaus[j] ? ld4(cand[j]) : ~seq[j];
        uint32_t nib = 0, dd[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool hit = cseq[j] == seq[j];
            dd[j] = hit ? uint32_t(p0 + j - cand[j]) : 0u;
            nib |= uint32_t(hit) << j;
        }
        reinterpret_cast<uint2*>(s_dist + 128 * buf)[lane] = make_uint2(dd[0] | (dd[1] << 16), dd[2] | (dd[3] << 16));
        uint32_t gw = nib << (4 * (lane & 7));        // OR over each group of 8 lanes (butterfly: all lanes in step)
        gw |= __shfl_xor_sync(B200_FULL, gw, 1); gw |= __shfl_xor_sync(B200_FULL, gw, 2); gw |= __shfl_xor_sync(B200_FULL, gw, 4);
        if ((lane & 7) == 0) s_mask[4 * buf + (lane >> 3)] = gw;
#if B200_V3_RUNS
        // dd[j] is 0 for a miss and >= 1 for a hit: position p continues a run when dd[p] == dd[p-1] != 0.  Position 0 of
        // a chunk always starts a run (the previous chunk's hits are another buffer's business).
        uint32_t pd = __shfl_up_sync(B200_FULL, dd[3], 1);
        if (lane == 0) pd = 0;
        uint32_t rs = nib;
        if (dd[0] && dd[0] == pd) rs &= ~1u;
        if (dd[1] && dd[1] == dd[0]) rs &= ~2u;
        if (dd[2] && dd[2] == dd[1]) rs &= ~4u;
        if (dd[3] && dd[3] == dd[2]) rs &= ~8u;
        uint32_t gr = rs << (4 * (lane & 7));
        gr |= __shfl_xor_sync(B200_FULL, gr, 1); gr |= __shfl_xor_sync(B200_FULL, gr, 2); gr |= __shfl_xor_sync(B200_FULL, gr, 4);
        if ((lane & 7) == 0) s_rmask[4 * buf + (lane >> 3)] = gr;
#endif
    };

    // ------------------------------------------------------------------ lay out the sequences of one chunk (warp E, or L)
    int op = 0, prev_end = 0; bool fail = false;
#if B200_V3_SPLIT
    auto emit = [&](int c) {                               // warp L: only the literal bytes are left to copy
        const int buf = c % NB;
        const int cnt = s_cnt[buf];
        op = s_state[4 * buf]; prev_end = s_state[4 * buf + 1]; fail = s_state[4 * buf + 2] != 0;
        if (fail || cnt == 0) return;
        uint4 r = make_uint4(0, 0, 0, 0);
        if (lane < cnt) r = s_rec4[32 * buf + lane];
        const int pe = int(r.x), lit = int(r.y), lpos = int(r.z);
        uint8_t* lo = dst + lpos;
        const int sn = min(lit, 16);
        const int mx = __reduce_max_sync(B200_FULL, sn);
        for (int t = 0; t < mx; t += 4) {
            if (t < sn) {
                const uint32_t v = ld4(pe + t);
                lo[t] = uint8_t(v);
                if (t + 1 < sn) lo[t + 1] = uint8_t(v >> 8);
                if (t + 2 < sn) lo[t + 2] = uint8_t(v >> 16);
                if (t + 3 < sn) lo[t + 3] = uint8_t(v >> 24);
            }
        }
        for (unsigned lm = __ballot_sync(B200_FULL, lit > 16); lm; lm &= lm - 1) {
            const int k = __ffs(lm) - 1;
            const int ka = __shfl_sync(B200_FULL, pe, k);
            const int kl = __shfl_sync(B200_FULL, lit, k);
            const int ko = __shfl_sync(B200_FULL, lpos, k);
            warp_copy(dst + ko + 16, src + ka + 16, kl - 16, lane);
        }
    };
#else
    auto emit = [&](int c) {
        const int buf = c % NB;
        const int cnt = s_cnt[buf];
        uint2 r = make_uint2(0, 4);
        if (lane < cnt) r = s_rec[32 * buf + lane];
        const int ms = int(r.x & 0xFFFFu), dist = int(r.x >> 16), ml = int(r.y);
        const int end = ms + ml;
        int pe = __shfl_up_sync(B200_FULL, end, 1);
        if (lane == 0) pe = prev_end;
        const int lit = lane < cnt ? ms - pe : 0, mcode = ml - 4;
        const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
        const int mhdr = mcode >= 15 ? (mcode - 15) / 255 + 1 : 0;
        const int size = lane < cnt ? 1 + lhdr + lit + 2 + mhdr : 0;
        int incl = size;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(B200_FULL, incl, d); if (lane >= d) incl += y; }
        const int total = __shfl_sync(B200_FULL, incl, 31);
        if (cnt > 0) prev_end = __shfl_sync(B200_FULL, end, cnt - 1);
        if (fail) return;
        if (uint32_t(op) + uint32_t(total) > uint32_t(cap)) { fail = true; return; }              // lz4.c:1085-1088, 1158
        const int o = op + incl - size;
        if (lane < cnt) {
            uint8_t* d = dst + o;
            d[0] = uint8_t((min(lit, 15) << 4) | min(mcode, 15));
            d += 1;
            if (lit >= 15) { int v = lit - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
            d += lit;
            d[0] = uint8_t(dist); d[1] = uint8_t(dist >> 8);                                     // LE16 offset (lz4.c:1133)
            d += 2;
            if (mcode >= 15) { int v = mcode - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
        }
        // Literal runs: every lane copies the first 16 bytes of its own run (the runs of one chunk lie within a
        // few lines of each other, so the 32 lanes' byte accesses coalesce); the rare longer runs are finished
        // by the whole warp.
        uint8_t* lo = dst + o + 1 + lhdr;
        const int sn = min(lit, 16);
        const int mx = __reduce_max_sync(B200_FULL, sn);
        for (int t = 0; t < mx; t += 4) {
            if (t < sn) {
                const uint32_t v = ld4(pe + t);
                lo[t] = uint8_t(v);
                if (t + 1 < sn) lo[t + 1] = uint8_t(v >> 8);
                if (t + 2 < sn) lo[t + 2] = uint8_t(v >> 16);
                if (t + 3 < sn) lo[t + 3] = uint8_t(v >> 24);
            }
        }
        for (unsigned lm = __ballot_sync(B200_FULL, lit > 16); lm; lm &= lm - 1) {
            const int k = __ffs(lm) - 1;
            const int ka = __shfl_sync(B200_FULL, pe, k);
            const int kl = __shfl_sync(B200_FULL, lit, k);
            const int ko = __shfl_sync(B200_FULL, o + 1 + lhdr, k);
            warp_copy(dst + ko + 16, src + ka + 16, kl - 16, lane);
        }
        op += total;
    };
#endif
    auto finish = [&]() {                                  // last literals (lz4.c:1266-1293)
        int ret = 0;
        if (!fail) {
            const int fin = prev_end;
            const int lit = n - fin;
            const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
            if (uint32_t(op) + 1u + uint32_t(lhdr) + uint32_t(lit) <= uint32_t(cap)) {
                if (lane == 0) dst[op] = uint8_t(min(lit, 15) << 4);
                op += 1;
                if (lhdr) { write_len_ext(dst + op, lit - 15, lhdr, lane); op += lhdr; }
                warp_copy(dst + op, src + fin, lit, lane);
                ret = op + lit;
            }
        }
        if (lane == 0) result[b] = ret;
    };

    // ------------------------------------------------------------------ the greedy walk of one chunk (warp P)
    int ip = 0, anchor = 0;
#if B200_V3_RUNS
#if B200_V3_SPLIT
    int pop = 0; bool pfail = false;                       // warp P's own output offset / overflow flag
#endif
    auto walk = [&](int c) {
        const int cp0 = 128 * c - int(ph), buf = c % NB;
        const bool inner = cp0 >= 4 && cp0 + 128 + 32 <= n;  // every measurement window of this chunk lies inside the block
        int k = 0;
        // true when chunk position e is a hit that continues a run (warp-uniform or per-lane e)
        auto continues = [&](int e) -> bool {
            if (e <= 0 || e >= 128) return false;
            return (((s_mask[4 * buf + (e >> 5)] & ~s_rmask[4 * buf + (e >> 5)]) >> (e & 31)) & 1u) != 0;
        };
        while (ip < cp0 + 128) {
            // ---- rank the run starts from ip on; if ip itself lies inside a run, from that run's start
            const int r0 = max(ip - cp0, 0);
            const bool mid0 = continues(r0);
            int start0 = r0;
            if (mid0) {
                #pragma unroll
                for (int kk = 3; kk >= 0; kk--) {
                    uint32_t m = s_rmask[4 * buf + kk];
                    const int hi = r0 - 32 * kk;             // keep positions below r0
                    if (hi <= 0) m = 0; else if (hi < 32) m &= (1u << hi) - 1u;
                    if (m && start0 == r0) start0 = 32 * kk + 31 - __clz(m);
                }
            }
            int nh = 0;
            #pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                uint32_t m = s_rmask[4 * buf + kk];
                const int lo = start0 - 32 * kk;
                if (lo >= 32) m = 0; else if (lo > 0) m &= 0xFFFFFFFFu << lo;
                const int rk = nh + __popc(m & ((1u << lane) - 1u));
                if (((m >> lane) & 1u) && rk < 32) s_hit[rk] = uint8_t(32 * kk + lane);
                nh += __popc(m);
            }
            __syncwarp();
            if (nh == 0) break;
            // ---- every lane measures one run start
            int ms = 0, ml = 0, back = 0, dist = 1; bool longer = false;
            const bool have = lane < nh;
            if (have) {
                ms = cp0 + s_hit[lane];
                dist = s_dist[128 * buf + (ms - cp0)];
                const int mc = ms - dist;
                const int lim = matchlimit - ms, capl = min(lim, 32);
                const uint32_t am = uint32_t(ms + 4) + ph, ac = uint32_t(mc + 4) + ph;     // +4: (pos - 4) never negative in the view
                const uint32_t lastw = (uint32_t(n - 1) + ph) >> 2;
                uint32_t wm[7], wc[7];
                if (inner) {
                    const uint32_t* pm = wsrc + (am >> 2) - 2;
                    const uint32_t* pc = wsrc + (mc >= 4 ? (ac >> 2) - 2 : 0u);
                    #pragma unroll
                    for (int t = 0; t < 7; t++) { wm[t] = pm[t]; wc[t] = pc[t]; }
                } else {
                    #pragma unroll
                    for (int t = 0; t < 7; t++) {
                        wm[t] = wsrc[min((am >> 2) - 2 + t, lastw)];
                        wc[t] = mc >= 4 ? wsrc[min((ac >> 2) - 2 + t, lastw)] : 0u;
                    }
                }
                const uint32_t sm = (am & 3u) * 8u, sc = (ac & 3u) * 8u;
                if (mc >= 4) {
                    const uint32_t x = __funnelshift_r(wm[0], wm[1], sm) ^ __funnelshift_r(wc[0], wc[1], sc);
                    back = x ? (__clz(x) >> 3) : 4;
                }
                ml = 4;
                if (mc >= 4) {
                    #pragma unroll
                    for (int t = 2; t < 6; t++) {
                        const uint32_t x = __funnelshift_r(wm[t], wm[t + 1], sm) ^ __funnelshift_r(wc[t], wc[t + 1], sc);
                        if (x) { ml += (__ffs(x) - 1) >> 3; goto measured; }
                        ml += 4;
                    }
                }
                while (ml < capl) {
                    const uint32_t x = ld4(ms + ml) ^ ld4(mc + ml);
                    if (x) { ml += (__ffs(x) - 1) >> 3; break; }
                    ml += 4;
                }
            measured:
                if (ml >= capl) { ml = capl; longer = capl < lim; }
            }
            // keys ascend (ranked positions); empty lanes sort to the back.  succ(v) = first lane whose run starts at or after v.
            const int key = have ? ms : 0x7FFFFFFF;
            const int key31 = __shfl_sync(B200_FULL, key, 31);
            auto succ = [&](int v) -> int {
                int q = 0;
                #pragma unroll
                for (int st = 16; st; st >>= 1) { const int pk = __shfl_sync(B200_FULL, key, q + st - 1); if (pk < v) q += st; }
                if (q == 31 && key31 < v) q = 32;
                return q;
            };
            int end = ms + ml;
            // where the chain goes after this lane's match: the run start at or after its end, or — when the end is a hit
            // inside a run — that run (the nearest run start before the end), entered in the middle
            const int nxt = succ(end) | (int(continues(end - cp0)) << 8);
            const unsigned hm = __ballot_sync(B200_FULL, have);
            const unsigned lm = __ballot_sync(B200_FULL, have && longer);
            unsigned sel = 0, msel = 0;
            int j = 0; bool jmid = mid0;
            while (j < 32 && ((hm >> j) & 1u)) {            // the greedy chain: one shuffle per selected sequence
                sel |= 1u << j;
                if (jmid) msel |= 1u << j;
                int step;
                if ((lm >> j) & 1u) {                        // 32 bytes matched and more to go: finish with the whole warp
                    const int jend = __shfl_sync(B200_FULL, end, j), jdist = __shfl_sync(B200_FULL, dist, j);
                    const int ext = match_extend(InGlobal{src}, jend, jend - jdist, matchlimit - jend, lane);
                    if (lane == j) { ml += ext; end += ext; }
                    step = succ(jend + ext) | (int(continues(jend + ext - cp0)) << 8);
                } else step = __shfl_sync(B200_FULL, nxt, j);
                const int nj = step & 0xFF;
                jmid = (step >> 8) != 0;
                if (jmid && nj == 32 && nh > 32) { j = 32; break; }     // the run in question was not ranked into this round
                j = jmid ? nj - 1 : nj;
            }
            {
                const unsigned below = sel & ((1u << lane) - 1u);
                int pend = __shfl_sync(B200_FULL, end, (31 - __clz(below)) & 31);     // end of the previous selected sequence
                if (!below) pend = anchor;
                const bool selme = ((sel >> lane) & 1u) != 0;
                int start = pend, len = 4;
                if (selme) {
                    if ((msel >> lane) & 1u) { start = pend; len = end - pend; }         // entered inside the run: no catch-up
                    else { const int bk = min(back, ms - pend); start = ms - bk; len = ml + bk; }
                }
#if B200_V3_SPLIT
                // lay the headers out here (lz4.c:1094-1100, 1133, 1184-1196); warp L copies the literal bytes later
                const int lit = selme ? start - pend : 0, mcode = len - 4;
                const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
                const int mhdr = (selme && mcode >= 15) ? (mcode - 15) / 255 + 1 : 0;
                const int size = selme ? 1 + lhdr + lit + 2 + mhdr : 0;
                int incl = size;
                #pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(B200_FULL, incl, d); if (lane >= d) incl += y; }
                const int total = __shfl_sync(B200_FULL, incl, 31);
                if (!pfail && uint32_t(pop) + uint32_t(total) > uint32_t(cap)) pfail = true;         // lz4.c:1085-1088, 1158
                if (!pfail) {
                    const int o = pop + incl - size;
                    if (selme) {
                        uint8_t* d = dst + o;
                        d[0] = uint8_t((min(lit, 15) << 4) | min(mcode, 15));
                        d += 1;
                        if (lit >= 15) { int v = lit - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
                        d += lit;
                        d[0] = uint8_t(dist); d[1] = uint8_t(dist >> 8);
                        d += 2;
                        if (mcode >= 15) { int v = mcode - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
                        s_rec4[32 * buf + k + __popc(below)] = make_uint4(uint32_t(pend), uint32_t(lit), uint32_t(o + 1 + lhdr), 0u);
                    }
                    pop += total;
                }
#else
                if (selme) s_rec[32 * buf + k + __popc(below)] = make_uint2(uint32_t(start) | (uint32_t(dist) << 16), uint32_t(len));
#endif
                k += __popc(sel);
                ip = anchor = __shfl_sync(B200_FULL, end, 31 - __clz(sel));
            }
            if (nh <= 32) break;                             // every run start of the chunk was in this round
            __syncwarp();                                    // s_hit is re-ranked from the new ip
        }
        if (lane == 0) {
            s_cnt[buf] = k;
#if B200_V3_SPLIT
            s_state[4 * buf] = pop; s_state[4 * buf + 1] = anchor; s_state[4 * buf + 2] = int(pfail);
#endif
        }
    };
#else
    auto walk = [&](int c) {
        const int cp0 = 128 * c - int(ph), buf = c % NB;
        const bool inner = cp0 >= 4 && cp0 + 128 + 32 <= n;  // every measurement window of this chunk lies inside the block
        int k = 0;
        const int r0 = max(ip - cp0, 0);
        int nh = 0;
        if (r0 < 128) {
            #pragma unroll
            for (int kk = 0; kk < 4; kk++) {                // rank the hits at or after ip: s_hit[rank] = position in chunk
                uint32_t m = s_mask[4 * buf + kk];
                const int lo = r0 - 32 * kk;
                if (lo >= 32) m = 0; else if (lo > 0) m &= 0xFFFFFFFFu << lo;
                if ((m >> lane) & 1u) s_hit[nh + __popc(m & ((1u << lane) - 1u))] = uint8_t(32 * kk + lane);
                nh += __popc(m);
            }
        }
        __syncwarp();
        for (int done = 0; done < nh && ip < cp0 + 128; done += 32) {
            const int idx = done + lane;
            int ms = 0, ml = 0, back = 0, dist = 1; bool longer = false;
            bool have = idx < nh;
            if (have) { ms = cp0 + s_hit[idx]; have = ms >= ip; }
            if (have) {
                dist = s_dist[128 * buf + (ms - cp0)];
                const int mc = ms - dist;
                const int lim = matchlimit - ms, capl = min(lim, 32);
                // One batch of loads covers the 4 bytes before and the 16 after the verified 4 on both sides (the
                // candidate side misses L1 as a rule: one L2 round trip here instead of one per 4 bytes).
                const uint32_t am = uint32_t(ms + 4) + ph, ac = uint32_t(mc + 4) + ph;     // +4: (pos - 4) never negative in the view
                const uint32_t lastw = (uint32_t(n - 1) + ph) >> 2;
                uint32_t wm[7], wc[7];
                if (inner) {                                     // the 28 bytes behind the hit are inside the block: no clamping
                    const uint32_t* pm = wsrc + (am >> 2) - 2;
                    const uint32_t* pc = wsrc + (mc >= 4 ? (ac >> 2) - 2 : 0u);
                    #pragma unroll
                    for (int t = 0; t < 7; t++) { wm[t] = pm[t]; wc[t] = pc[t]; }
                } else {
                    #pragma unroll
                    for (int t = 0; t < 7; t++) {
                        wm[t] = wsrc[min((am >> 2) - 2 + t, lastw)];
                        wc[t] = mc >= 4 ? wsrc[min((ac >> 2) - 2 + t, lastw)] : 0u;
                    }
                }
                const uint32_t sm = (am & 3u) * 8u, sc = (ac & 3u) * 8u;
                if (mc >= 4) {
                    const uint32_t x = __funnelshift_r(wm[0], wm[1], sm) ^ __funnelshift_r(wc[0], wc[1], sc);
                    back = x ? (__clz(x) >> 3) : 4;
                }
                ml = 4;
                if (mc >= 4) {
                    #pragma unroll
                    for (int t = 2; t < 6; t++) {
                        const uint32_t x = __funnelshift_r(wm[t], wm[t + 1], sm) ^ __funnelshift_r(wc[t], wc[t + 1], sc);
                        if (x) { ml += (__ffs(x) - 1) >> 3; goto measured; }
                        ml += 4;
                    }
                }
                while (ml < capl) {
                    const uint32_t x = ld4(ms + ml) ^ ld4(mc + ml);
                    if (x) { ml += (__ffs(x) - 1) >> 3; break; }
                    ml += 4;
                }
            measured:
                if (ml >= capl) { ml = capl; longer = capl < lim; }
            }
            // Hit positions are ranked, so the lanes' keys ascend: stale hits (before ip) sort to the front, empty
            // lanes to the back.  succ(v) = first lane whose hit starts at or after v.
            const int key = have ? ms : (idx < nh ? -1 : 0x7FFFFFFF);
            const int key31 = __shfl_sync(B200_FULL, key, 31);
            auto succ = [&](int v) -> int {
                int q = 0;
                #pragma unroll
                for (int st = 16; st; st >>= 1) { const int pk = __shfl_sync(B200_FULL, key, q + st - 1); if (pk < v) q += st; }
                if (q == 31 && key31 < v) q = 32;
                return q;
            };
            int end = ms + ml;
            const int nxt = succ(end);
            const unsigned hm = __ballot_sync(B200_FULL, have);
            const unsigned lm = __ballot_sync(B200_FULL, have && longer);
            unsigned sel = 0;
            int j = hm ? __ffs(hm) - 1 : 32;
            while (j < 32 && ((hm >> j) & 1u)) {            // the greedy chain: one shuffle per selected sequence
                sel |= 1u << j;
                if ((lm >> j) & 1u) {                        // 32 bytes matched and more to go: finish with the whole warp
                    const int jend = __shfl_sync(B200_FULL, end, j), jdist = __shfl_sync(B200_FULL, dist, j);
                    const int ext = match_extend(InGlobal{src}, jend, jend - jdist, matchlimit - jend, lane);
                    if (lane == j) { ml += ext; end += ext; }
                    j = succ(jend + ext);
                } else j = __shfl_sync(B200_FULL, nxt, j);
            }
            if (sel) {
                const unsigned below = sel & ((1u << lane) - 1u);
                int pend = __shfl_sync(B200_FULL, end, (31 - __clz(below)) & 31);     // end of the previous selected sequence
                if (!below) pend = anchor;
                if ((sel >> lane) & 1u) {
                    const int bk = min(back, ms - pend);
                    s_rec[32 * buf + k + __popc(below)] = make_uint2(uint32_t(ms - bk) | (uint32_t(dist) << 16), uint32_t(ml + bk));
                }
                k += __popc(sel);
                ip = anchor = __shfl_sync(B200_FULL, end, 31 - __clz(sel));
            }
        }
        if (lane == 0) s_cnt[buf] = k;
    };

#endif

    if (role == 0) {
        for (int i = lane; i < TABLE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        if (NW == 2) {
            for (int i = 0; i < nchunks + LAG; i++) {
                // chunk i's buffer is free: its previous tenant (chunk i-NB) was laid out in iteration i-1
                if (i < nchunks) { lookup(i); bar_arrive(BAR_FULL + i % NB); }
                if (i >= LAG) { bar_wait(BAR_WALKED + (i - LAG) % NB); emit(i - LAG); }
            }
            finish();
        } else {
            for (int c = 0; c < nchunks; c++) {
                if (c >= NB) bar_wait(BAR_DFREE + c % NB);                      // warp P is done with chunk c-NB's distances
                lookup(c);
                bar_arrive(BAR_FULL + c % NB);
            }
        }
    } else if (role == 1) {
        for (int c = 0; c < nchunks; c++) {
            bar_wait(BAR_FULL + c % NB);
            if (NW == 3 && c >= NB) bar_wait(BAR_RFREE + c % NB);               // warp E is done with chunk c-NB's records
            walk(c);
            bar_arrive(BAR_WALKED + c % NB);
            if (NW == 3 && c + NB < nchunks) bar_arrive(BAR_DFREE + c % NB);
        }
    } else {
        for (int c = 0; c < nchunks; c++) {
            bar_wait(BAR_WALKED + c % NB);
            emit(c);
            if (c + NB < nchunks) bar_arrive(BAR_RFREE + c % NB);
        }
        finish();
    }
}

#ifndef B200_HOST_SIM
template <int HASH_LOG, bool SPARSE>
static cudaError_t launch_v3(const BatchArgs& a, cudaStream_t st)
{
    const size_t smem = (2u << HASH_LOG) + B200_V3_NB * (256 + 256 + 16) + 16 + 128 + (B200_V3_RUNS ? B200_V3_NB * 16 : 0) + (B200_V3_SPLIT ? B200_V3_NB * (256 + 16) : 0);
    auto k = lz4_compress_fast3_kernel<HASH_LOG, SPARSE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<(unsigned)a.n, 32 * B200_V3_WARPS, smem, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap,
                                                        a.result, (uint32_t)a.n);
    return cudaGetLastError();
}
#endif

// ---------------------------------------------------------------------------------------------
// Three-kernel pipeline (algo 4): the three phases of the decoupled parser as three launches over a
// sub-batch, each shaped for what it does, with the per-position state in a global scratch arena:
//   K1 lookup   one warp per block, hash table in shared memory (13 warps/SM): phase AB only — distances
//               (u16 per position) and hit masks go to the arena, fully coalesced.
//   K2 walk     ONE THREAD per block (32 blocks per warp, no shared memory, 64 warps/SM): the serial greedy
//               walk is scalar work, so it runs as scalar work at 32x the multiplicity; match extension is
//               per-lane up to 32 bytes, longer matches are finished cooperatively by the whole warp.
//               Output: one 8-byte record per sequence.
//   K3 layout   one warp per block (64 warps/SM): prefix sums of sequence sizes, 32 tokens/offsets per
//               instruction, cooperative literal copies, last literals, result.
// Same parse and same bytes as algos 2/3.  The arena costs extra memory traffic (it does not fit L2);
// DESIGN.md discusses that trade.
static constexpr int K4_DIST_STRIDE = 65536 + 128;     // u16 entries per block, indexed by aligned-view byte
static constexpr int K4_MASK_STRIDE = 2064;            // u32 words per block
static constexpr int K4_REC_STRIDE  = 16400;           // uint2 records per block (>= 65536/4 + slack)

template <int HASH_LOG, bool SPARSE>
__global__ void __launch_bounds__(32)
lz4c4_lookup_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                    const int32_t* __restrict__ src_len, uint32_t first, uint32_t nsb,
                    uint16_t* __restrict__ g_dist, uint32_t* __restrict__ g_mask)
{
    B200_DYN_SMEM(smem_raw, 128);
    uint16_t* table = reinterpret_cast<uint16_t*>(smem_raw);
    const uint32_t sb = blockIdx.x;
    if (sb >= nsb) return;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[first + sb];
    const int n = src_len[first + sb];
    if (n < 13 || n >= 65536 + 11) return;                        // nothing to look up (K3 handles these sizes)
    for (int i = lane; i < (2 << HASH_LOG) / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    const uint32_t ph = uint32_t(reinterpret_cast<uintptr_t>(src)) & 3u;
    const uint32_t* __restrict__ wsrc = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) - ph);
    const int mflimit = n - 12;
    const int nchunks = (mflimit + int(ph)) / 128 + 1;
    uint2* dout = reinterpret_cast<uint2*>(g_dist + size_t(sb) * K4_DIST_STRIDE);
    uint32_t* mout = g_mask + size_t(sb) * K4_MASK_STRIDE;
    for (int c = 0; c < nchunks; c++) {
        const int cp0 = 128 * c - int(ph);
        if (lane < 2) {
            const int pfq = cp0 + 512 + lane * 128;
            if (pfq < n) B200_PREFETCH_L2(src + pfq);
        }
        const int p0 = cp0 + 4 * lane;
        uint32_t w0 = 0, w1 = 0;
        if (p0 + 3 >= 0 && p0 <= mflimit) { w0 = wsrc[32 * c + lane]; w1 = wsrc[32 * c + lane + 1]; }
        uint32_t seq[4], h[4]; int cand[4]; bool plaus[4];
        seq[0] = w0; seq[1] = __funnelshift_r(w0, w1, 8); seq[2] = __funnelshift_r(w0, w1, 16); seq[3] = __funnelshift_r(w0, w1, 24);
        #pragma unroll
        for (int j = 0; j < 4; j++) { h[j] = (seq[j] * 2654435761u) >> (32 - HASH_LOG); cand[j] = table[h[j]]; }
        __syncwarp();
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p = p0 + j;
            const bool valid = p >= 0 && p <= mflimit;
            if (valid && (!SPARSE || j == 0)) table[h[j]] = uint16_t(p);
            plaus[j] = valid && cand[j] < p;
        }
        uint32_t cseq[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t v = ~seq[j];
            if (plaus[j]) { const uint32_t a = uint32_t(cand[j]) + ph; const uint32_t* w = wsrc + (a >> 2); v = __funnelshift_r(w[0], w[1], (a & 3u) * 8u); }
            cseq[j] = v;
        }
        uint32_t nib = 0, dd[4];
        #pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool hit = cseq[j] == seq[j];
            dd[j] = hit ? uint32_t(p0 + j - cand[j]) : 0u;
            nib |= uint32_t(hit) << j;
        }
        dout[32 * c + lane] = make_uint2(dd[0] | (dd[1] << 16), dd[2] | (dd[3] << 16));
        uint32_t gw = nib << (4 * (lane & 7));
        gw |= __shfl_xor_sync(B200_FULL, gw, 1); gw |= __shfl_xor_sync(B200_FULL, gw, 2); gw |= __shfl_xor_sync(B200_FULL, gw, 4);
        if ((lane & 7) == 0) mout[4 * c + (lane >> 3)] = gw;
        __syncwarp();                                              // this chunk's inserts precede the next chunk's lookups
    }
}

// K2: one thread per block.  Coordinates are bytes of the aligned view (a = position + ph).
__global__ void __launch_bounds__(128)
lz4c4_walk_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                  const int32_t* __restrict__ src_len, uint32_t first, uint32_t nsb,
                  const uint16_t* __restrict__ g_dist, const uint32_t* __restrict__ g_mask,
                  uint2* __restrict__ g_rec, int32_t* __restrict__ g_cnt)
{
    const uint32_t sb = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = lane_id();
    const bool live = sb < nsb;
    const uint8_t* src = live ? src_base + src_off[first + sb] : src_base;
    const int n = live ? src_len[first + sb] : 0;
    const bool work = live && n >= 13 && n < 65536 + 11;
    const int ph = int(reinterpret_cast<uintptr_t>(src) & 3u);
    const int mflimit = n - 12, matchlimit = n - 5;
    const uint16_t* ds = g_dist + size_t(sb) * K4_DIST_STRIDE;
    const uint32_t* mk = g_mask + size_t(sb) * K4_MASK_STRIDE;
    uint2* rec = g_rec + size_t(sb) * K4_REC_STRIDE;
    const int nwords = work ? ((mflimit + ph) >> 5) + 1 : 0;
    int w = -1; uint32_t m = 0;
    int ip = 0, anchor = 0, nrec = 0;        // positions (not view bytes)

    for (;;) {
        // ---- each lane advances to its next hit at or after ip
        bool have = false; int ms = 0;
        while (work) {
            if (m == 0) {
                w++;
                if (w >= nwords) break;
                m = mk[w];
                const int lo = ip + ph - 32 * w;                   // first still-eligible bit of this word
                if (lo >= 32) { m = 0; continue; }
                if (lo > 0) m &= 0xFFFFFFFFu << lo;
                continue;
            }
            const int q = __ffs(m) - 1;
            m &= m - 1;
            const int p = 32 * w + q - ph;
            if (p < ip) continue;
            ms = p; have = true;
            break;
        }
        if (__ballot_sync(B200_FULL, have) == 0) break;            // every lane of the warp is out of hits
        int ml = 0, dist = 0; bool longer = false;
        if (have) {
            dist = ds[ms + ph];
            const int mc = ms - dist;
            // catch-up (lz4.c:1080), at most 8 bytes
            int back = 0;
            const int backroom = min(min(ms - anchor, mc), 8);
            while (back < backroom && src[ms - 1 - back] == src[mc - 1 - back]) back++;
            // match body (lz4.c:1153): bytes are known equal for 4; compare words up to a 36-byte cap here
            ml = 4;
            const int lim = matchlimit - ms;
            const int cap = min(lim, 36);
            while (ml < cap) {
                const uint32_t x = load_u32_unaligned(src + ms + ml) ^ load_u32_unaligned(src + mc + ml);
                if (x) { ml += (__ffs(x) - 1) >> 3; break; }
                ml += 4;
            }
            if (ml >= cap) { ml = cap; longer = cap < lim; }
            ms -= back; ml += back;
        }
        // ---- matches that ran into the cap are finished by the whole warp, 128 bytes per round
        unsigned todo = __ballot_sync(B200_FULL, longer);
        while (todo) {
            const int l = __ffs(todo) - 1; todo &= todo - 1;
            const unsigned long long sp = __shfl_sync(B200_FULL, (unsigned long long)reinterpret_cast<uintptr_t>(src), l);
            const int e_ms = __shfl_sync(B200_FULL, ms, l), e_ml = __shfl_sync(B200_FULL, ml, l);
            const int e_dist = __shfl_sync(B200_FULL, dist, l), e_lim = __shfl_sync(B200_FULL, matchlimit, l);
            const uint8_t* s2 = reinterpret_cast<const uint8_t*>(sp);
            const int more = match_extend(InGlobal{s2}, e_ms + e_ml, e_ms + e_ml - e_dist, e_lim - (e_ms + e_ml), lane);
            if (lane == l) ml += more;
        }
        if (have) {
            rec[nrec++] = make_uint2(uint32_t(ms) | (uint32_t(dist) << 16), uint32_t(ml));
            ip = anchor = ms + ml;
            const int lo = ip + ph - 32 * w;                       // drop the hits this match covered
            if (lo >= 32) m = 0; else if (lo > 0) m &= 0xFFFFFFFFu << lo;
        }
    }
    if (live) g_cnt[sb] = nrec;
}

// K3: one warp per block writes the LZ4 stream from the records
__global__ void __launch_bounds__(128)
lz4c4_layout_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                    const int32_t* __restrict__ src_len,
                    uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                    const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t first, uint32_t nsb,
                    const uint2* __restrict__ g_rec, const int32_t* __restrict__ g_cnt)
{
    const uint32_t sb = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (sb >= nsb) return;
    const uint32_t b = first + sb;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b];
    const int cap = dst_cap[b];
    int ret = 0;
    if (n < 0 || n >= 65536 + 11) goto done;
    if (n == 0) { if (cap >= 1) { if (lane == 0) dst[0] = 0; ret = 1; } goto done; }
    {
        const int cnt = n >= 13 ? g_cnt[sb] : 0;
        const uint2* rec = g_rec + size_t(sb) * K4_REC_STRIDE;
        int op = 0, anchor = 0;
        for (int base = 0; base < cnt; base += 32) {
            const int k = base + lane;
            const bool on = k < cnt;
            uint2 r = make_uint2(0, 4);
            if (on) r = rec[k];
            const int ms = int(r.x & 0xFFFFu), dist = int(r.x >> 16), ml = int(r.y);
            const int end = ms + ml;
            int prev_end = __shfl_up_sync(B200_FULL, end, 1);
            if (lane == 0) prev_end = anchor;
            const int lit = on ? ms - prev_end : 0, mcode = ml - 4;
            const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
            const int mhdr = mcode >= 15 ? (mcode - 15) / 255 + 1 : 0;
            const int size = on ? 1 + lhdr + lit + 2 + mhdr : 0;
            int incl = size;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(B200_FULL, incl, d); if (lane >= d) incl += y; }
            const int total = __shfl_sync(B200_FULL, incl, 31);
            if (uint32_t(op) + uint32_t(total) > uint32_t(cap)) goto done;                        // lz4.c:1085-1088, 1158
            const int o = op + incl - size;
            if (on) {
                uint8_t* d = dst + o;
                d[0] = uint8_t((min(lit, 15) << 4) | min(mcode, 15));
                d += 1;
                if (lit >= 15) { int v = lit - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
                d += lit;
                d[0] = uint8_t(dist); d[1] = uint8_t(dist >> 8);
                d += 2;
                if (mcode >= 15) { int v = mcode - 15; for (; v >= 255; v -= 255) *d++ = 255; *d++ = uint8_t(v); }
            }
            const int nk = min(32, cnt - base);
            for (int kk = 0; kk < nk; kk++) {
                const int ka = __shfl_sync(B200_FULL, prev_end, kk);
                const int kl = __shfl_sync(B200_FULL, lit, kk);
                const int ko = __shfl_sync(B200_FULL, o + 1 + lhdr, kk);
                warp_copy(dst + ko, src + ka, kl, lane);
            }
            op += total;
            anchor = __shfl_sync(B200_FULL, end, nk - 1);
        }
        {   // last literals (lz4.c:1266-1293)
            const int lit = n - anchor;
            const int lhdr = lit >= 15 ? (lit - 15) / 255 + 1 : 0;
            if (uint32_t(op) + 1u + uint32_t(lhdr) + uint32_t(lit) > uint32_t(cap)) goto done;
            if (lane == 0) dst[op] = uint8_t(min(lit, 15) << 4);
            op += 1;
            if (lhdr) { write_len_ext(dst + op, lit - 15, lhdr, lane); op += lhdr; }
            warp_copy(dst + op, src + anchor, lit, lane);
            ret = op + lit;
        }
    }
done:
    if (lane == 0) result[b] = ret;
}

// scratch arena: one per (calling thread, stream) so concurrent pipelines never share it
#ifndef B200_HOST_SIM          // host side: arenas, launchers, knobs
struct K4Arena { cudaStream_t st; int device; uint32_t blocks; uint16_t* dist; uint32_t* mask; uint2* rec; int32_t* cnt; };
static thread_local K4Arena t_arenas[8];
static thread_local int t_narenas = 0;
extern "C" { int b200lz4_compress_subbatch = 16384; }   // blocks per K1/K2/K3 round (arena = 270 KB per block)

static cudaError_t k4_arena(cudaStream_t st, uint32_t blocks, K4Arena** out)
{
    int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    K4Arena* a = nullptr;
    for (int i = 0; i < t_narenas; i++) if (t_arenas[i].st == st && t_arenas[i].device == dev) a = &t_arenas[i];
    if (!a) {
        if (t_narenas == 8) { a = &t_arenas[0]; cudaFree(a->dist); cudaFree(a->mask); cudaFree(a->rec); cudaFree(a->cnt); }
        else a = &t_arenas[t_narenas++];
        *a = K4Arena{st, dev, 0, nullptr, nullptr, nullptr, nullptr};
    }
    if (a->blocks < blocks) {
        if (a->dist) { cudaFree(a->dist); cudaFree(a->mask); cudaFree(a->rec); cudaFree(a->cnt); }
        a->blocks = 0;
        if ((e = cudaMalloc(&a->dist, size_t(blocks) * K4_DIST_STRIDE * 2)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&a->mask, size_t(blocks) * K4_MASK_STRIDE * 4)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&a->rec, size_t(blocks) * K4_REC_STRIDE * 8)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&a->cnt, size_t(blocks) * 4)) != cudaSuccess) return e;
        a->blocks = blocks;
    }
    *out = a;
    return cudaSuccess;
}

template <int HASH_LOG, bool SPARSE>
static cudaError_t launch_v4(const BatchArgs& a, cudaStream_t st)
{
    const uint32_t sbmax = (uint32_t)std::min<size_t>(a.n, (size_t)std::max(b200lz4_compress_subbatch, 32));
    K4Arena* ar; cudaError_t e = k4_arena(st, sbmax, &ar); if (e != cudaSuccess) return e;
    auto k1 = lz4c4_lookup_kernel<HASH_LOG, SPARSE>;
    const size_t smem = 2u << HASH_LOG;
    if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
    cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    for (size_t first = 0; first < a.n; first += sbmax) {
        const uint32_t nsb = (uint32_t)std::min<size_t>(sbmax, a.n - first);
        k1<<<nsb, 32, smem, st>>>(a.src_base, a.src_off, a.src_len, (uint32_t)first, nsb, ar->dist, ar->mask);
        lz4c4_walk_kernel<<<(nsb + 127) / 128, 128, 0, st>>>(a.src_base, a.src_off, a.src_len, (uint32_t)first, nsb, ar->dist, ar->mask, ar->rec, ar->cnt);
        lz4c4_layout_kernel<<<(nsb + 3) / 4, 128, 0, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result,
                                                          (uint32_t)first, nsb, ar->rec, ar->cnt);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int HASH_LOG, int NB, int NW>
static cudaError_t launch_v5(const BatchArgs& a, cudaStream_t st)
{
    constexpr int S = 2;                                                     // sub-rounds of 128 positions per chunk
    using LY = WideLayout<S, NB, NW>;
    const size_t smem = LY::smem(HASH_LOG);
    constexpr int FIT = 233472 / (int(LY::smem(HASH_LOG)) + 1024);          // CTAs per SM that shared memory allows
    constexpr int MINB = FIT < 16 ? FIT : 16;
    auto k = lz4_compress_wide_kernel<HASH_LOG, S, NB, NW, MINB>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<(unsigned)a.n, 32 * NW, smem, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, (uint32_t)a.n);
    return cudaGetLastError();
}

template <int HASH_LOG, bool U16, bool STAGE>
static cudaError_t launch_variant(const BatchArgs& a, cudaStream_t st)
{
    const size_t smem = ((U16 ? 2u : 4u) << HASH_LOG) + (STAGE ? 16 + STAGE_BYTES : 0);
    auto k = lz4_compress_fast_kernel<HASH_LOG, U16, STAGE>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // occupancy here is bounded by shared bytes per warp: take the largest carve-out
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<(unsigned)a.n, 32, smem, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap,
                                       a.result, (uint32_t)a.n);
    return cudaGetLastError();
}

// tuning knobs (not part of the public header; tools/ and bench.py may set them through ctypes)
extern "C" {
int b200lz4_compress_hash_log = 13;   // 13 = the reference's table size for <64 KiB blocks (lz4.c:756-762)
int b200lz4_compress_stage = 0;       // 1 = stage <=64 KiB blocks in shared memory via TMA (v1 parser only)
int b200lz4_compress_sparse = 0;      // 1 = publish one position in four (pairs with hash_log 12: the 'fast' operating point)
int b200lz4_compress_algo = 5;        // 3 = decoupled, two-warp pipeline (default); 2 = decoupled, one warp; 1 = coupled warp parser
int b200lz4_compress_wide = 322;      // algo 5: 100 * warps + 10 * sub-rounds per chunk + chunk buffers
}

cudaError_t launch_compress_fast(const BatchArgs& a, int max_src_len, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    const bool u16 = max_src_len > 0 && max_src_len <= 65536;
    if (b200lz4_compress_algo == 5 && u16) {
        switch (b200lz4_compress_wide) {               // 100 * warps + 20 + chunk buffers
        case 322: return launch_v5<13, 2, 3>(a, st);
        default: return launch_v5<13, 2, 2>(a, st);
        }
    }
    if (b200lz4_compress_algo == 4 && !b200lz4_compress_stage && u16) {
        if (b200lz4_compress_hash_log == 12) return b200lz4_compress_sparse ? launch_v4<12, true>(a, st) : launch_v4<12, false>(a, st);
        return b200lz4_compress_sparse ? launch_v4<13, true>(a, st) : launch_v4<13, false>(a, st);
    }
    if ((b200lz4_compress_algo == 3 || b200lz4_compress_algo == 4) && !b200lz4_compress_stage) {
        if (!u16) return launch_v2<12, false>(a, st);          // blocks > 64 KiB: one-warp decoupled parser with the 32-bit table
        if (b200lz4_compress_hash_log == 12) return b200lz4_compress_sparse ? launch_v3<12, true>(a, st) : launch_v3<12, false>(a, st);
        return b200lz4_compress_sparse ? launch_v3<13, true>(a, st) : launch_v3<13, false>(a, st);
    }
    if (b200lz4_compress_algo == 2 && !b200lz4_compress_stage) {
        if (!u16) return launch_v2<12, false>(a, st);
        if (b200lz4_compress_hash_log == 12) return launch_v2<12, true>(a, st);
        return launch_v2<13, true>(a, st);
    }
    if (u16) {
        if (b200lz4_compress_stage) {
            if (b200lz4_compress_hash_log == 12) return launch_variant<12, true, true>(a, st);
            return launch_variant<13, true, true>(a, st);
        }
        if (b200lz4_compress_hash_log == 12) return launch_variant<12, true, false>(a, st);
        if (b200lz4_compress_hash_log == 11) return launch_variant<11, true, false>(a, st);
        return launch_variant<13, true, false>(a, st);
    }
    return launch_variant<12, false, false>(a, st);   // 4096 x u32 = 16 KiB, the reference's byU32 table (lz4.c:1356)
}

#endif

} // namespace b200
