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
#include <type_traits>

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
                if (qq < n) asm volatile("prefetch.global.L2 [%0];" :: "l"(__cvta_generic_to_global(gsrc + qq)));
                pf += 4096;
            }
        }
        // ---- A: probe 32 positions; the pending sequence's literal bytes are fetched alongside
        uint32_t litv = 0;
        if (have && lane < q.ms - q.anchor && q.ms - q.anchor <= 32) litv = in.ld1(q.anchor + lane);
        if (std::is_same<In, InGlobal>::value && lane == 0 && ip + 160 < n)     // next line of the forward stream -> L1
            asm volatile("prefetch.global.L1 [%0];" :: "l"(__cvta_generic_to_global(gsrc + ip + 128)));
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
    extern __shared__ __align__(128) uint8_t smem_raw[];
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
        const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
        const int bulk = aligned ? (n & ~15) : 0;
        if (lane == 0 && bulk) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bulk) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"((uint32_t)__cvta_generic_to_shared(stage)), "l"(src), "r"(bulk), "r"(bar) : "memory");
        }
        for (int i = bulk + lane; i < n; i += 32) stage[i] = src[i];
        __syncwarp();
        if (bulk) {
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar) : "memory");
            } while (!ok);
        }
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
// Decoupled parser ("v2"): candidate lookup + verification run AHEAD of the serial greedy parse.
//
// The v1 kernel above discovers one match, extends it, writes it, and only then knows where to probe
// next: three dependent L2 round trips per ~27 input bytes with 13 warps/SM to hide them.  Here the
// block is processed in windows of W positions and each window in two phases:
//   AB (throughput, lane-parallel, no decisions): every position is hashed, looked up and INSERTED in
//      order (32 per step, lookups of a step before its inserts), every candidate is verified and each
//      lane measures its own match (forward up to CAP bytes, backward up to 4) — results go to a small
//      per-warp array in shared memory plus a hit bitmap held in registers.  Four steps are in flight
//      at once so their loads overlap; nothing here depends on the parse.
//   C  (serial, cheap): the greedy walk only reads that array: next hit at or after ip from the bitmap
//      (ballot/ffs/shuffle), distance + lengths by one shared-memory read, cooperative extension only
//      for matches that hit CAP, then the usual sequence emission.
// Inserting every position (not only those before a match start) costs no ratio: on the reference's
// own generator the parse is slightly denser than lz4's (1.629 vs 1.612 at P=0.50).
template <int HASH_LOG, bool U16>
__global__ void __launch_bounds__(32)
lz4_compress_fast2_kernel(const uint8_t* __restrict__ src_base, const uint64_t* __restrict__ src_off,
                          const int32_t* __restrict__ src_len,
                          uint8_t* __restrict__ dst_base, const uint64_t* __restrict__ dst_off,
                          const int32_t* __restrict__ dst_cap, int32_t* __restrict__ result, uint32_t nblocks)
{
    using Entry = typename std::conditional<U16, uint16_t, uint32_t>::type;
    constexpr int W = 256, STEPS = W / 32, G = 4, CAP = 32;    // 1 KiB of window state: 12 warps/SM next to a 16 KiB table
    constexpr int TABLE_BYTES = int(sizeof(Entry) << HASH_LOG);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    Entry* table = reinterpret_cast<Entry*>(smem_raw);
    uint16_t* s_dist = reinterpret_cast<uint16_t*>(smem_raw + TABLE_BYTES);        // [W] match distance, 0 = no match here
    uint16_t* s_mlb = s_dist + W;                                                  // [W] forward length | backward length << 8

    const uint32_t b = blockIdx.x;
    if (b >= nblocks) return;
    const int lane = lane_id();
    const uint8_t* __restrict__ src = src_base + src_off[b];
    uint8_t* __restrict__ dst = dst_base + dst_off[b];
    const int n = src_len[b];
    const int cap = dst_cap[b];
    int ret = 0;

    if (n < 0 || n > 0x7E000000) goto done;                       // lz4.c:1324
    if (U16 && n >= 65536 + 11) goto done;                         // lz4.c:973
    if (n == 0) { if (cap >= 1) { if (lane == 0) dst[0] = 0; ret = 1; } goto done; }
    {
        for (int i = lane; i < TABLE_BYTES / 16; i += 32) reinterpret_cast<uint4*>(table)[i] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        const InGlobal in{src};
        const int mflimit = n - 12, matchlimit = n - 5;
        int op = 0, anchor = 0, ip = 0;
        bool have = false; Seq q = {0, 0, 0, 0};

        for (int wbase = 0; wbase <= mflimit; wbase += W) {
            if (lane < W / 128) {                                  // the window after next on its way to L2
                const int pfq = wbase + 4 * W + lane * 128;
                if (pfq < n) asm volatile("prefetch.global.L2 [%0];" :: "l"(__cvta_generic_to_global(src + pfq)));
            }
            // ---------------- phase AB
            uint32_t hitword = 0;                                  // lane s keeps the hit mask of step s
            #pragma unroll 1
            for (int g = 0; g < STEPS; g += G) {
                uint32_t seq[G]; int cand[G]; bool plaus[G]; uint32_t cseq[G];
                #pragma unroll
                for (int k = 0; k < G; k++) {
                    const int p = wbase + (g + k) * 32 + lane;
                    seq[k] = in.ld4(min(p, mflimit));
                }
                #pragma unroll
                for (int k = 0; k < G; k++) {                      // table traffic in position order
                    const int p = wbase + (g + k) * 32 + lane;
                    const uint32_t h = (seq[k] * 2654435761u) >> (32 - HASH_LOG);
                    cand[k] = table[h];
                    if (p <= mflimit) table[h] = Entry(p);
                    plaus[k] = p <= mflimit && cand[k] < p && (U16 || p - cand[k] <= 65535);
                }
                #pragma unroll
                for (int k = 0; k < G; k++) {
                    const int p = wbase + (g + k) * 32 + lane;
                    cseq[k] = in.ld4(plaus[k] ? cand[k] : min(p, mflimit));
                }
                #pragma unroll
                for (int k = 0; k < G; k++) {
                    const int p = wbase + (g + k) * 32 + lane;
                    const bool hit = plaus[k] && cseq[k] == seq[k];
                    int ml = 0, back = 0;
                    if (hit) {
                        const int c = cand[k];
                        const int maxlen = min(matchlimit - p, CAP);
                        ml = 4;
                        while (ml < maxlen) {
                            const uint32_t x = in.ld4(p + ml) ^ in.ld4(c + ml);
                            if (x) { ml += (__ffs(x) - 1) >> 3; break; }
                            ml += 4;
                        }
                        ml = min(ml, maxlen);
                        if (c >= 4) {                               // catch-up potential: equal bytes just before both
                            const uint32_t x = in.ld4(p - 4) ^ in.ld4(c - 4);
                            back = x ? (__clz(x) >> 3) : 4;
                        }
                    }
                    const int idx = (g + k) * 32 + lane;
                    s_dist[idx] = hit ? uint16_t(p - cand[k]) : uint16_t(0);
                    s_mlb[idx] = uint16_t(ml | (back << 8));
                    const uint32_t m = __ballot_sync(B200_FULL, hit);
                    if (lane == g + k) hitword = m;
                }
            }
            __syncwarp();
            // ---------------- phase C: greedy walk over this window's hits
            const int wend = min(wbase + W, mflimit + 1);
            if (ip < wbase) ip = wbase;
            while (ip < wend) {
                const int r = ip - wbase;
                uint32_t mine = 0;
                if (lane < STEPS) {
                    if (lane == (r >> 5)) mine = hitword & (0xFFFFFFFFu << (r & 31));
                    else if (lane > (r >> 5)) mine = hitword;
                }
                const uint32_t any = __ballot_sync(B200_FULL, mine != 0);
                if (any == 0) break;
                const int wi = __ffs(any) - 1;
                const uint32_t wv = __shfl_sync(B200_FULL, mine, wi);
                const int qi = wi * 32 + __ffs(wv) - 1;             // window-relative position of the next match
                int ms = wbase + qi;
                const int dist = s_dist[qi];
                const uint32_t mlb = s_mlb[qi];
                int ml = int(mlb & 0xFF);
                const int back = min(int(mlb >> 8), ms - anchor);
                if (ml == CAP && ms + ml < matchlimit)
                    ml += match_extend(in, ms + ml, ms - dist + ml, matchlimit - (ms + ml), lane);
                // write the PREVIOUS sequence now (its literal bytes were requested one iteration ago)
                uint32_t litv = 0;
                if (have) {
                    if (lane < q.ms - q.anchor && q.ms - q.anchor <= 32) litv = in.ld1(q.anchor + lane);
                    if (!emit_sequence(in, q, litv, dst, op, cap, lane)) goto done;
                }
                ms -= back; ml += back;
                q.anchor = anchor; q.ms = ms; q.off = dist; q.ml = ml; have = true;
                ip = anchor = ms + ml;
            }
        }
        if (have) {
            uint32_t litv = 0;
            if (lane < q.ms - q.anchor && q.ms - q.anchor <= 32) litv = in.ld1(q.anchor + lane);
            if (!emit_sequence(in, q, litv, dst, op, cap, lane)) goto done;
        }
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

template <int HASH_LOG, bool U16>
static cudaError_t launch_v2(const BatchArgs& a, cudaStream_t st)
{
    const size_t smem = ((U16 ? 2u : 4u) << HASH_LOG) + 2 * 256 * sizeof(uint16_t);
    auto k = lz4_compress_fast2_kernel<HASH_LOG, U16>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    k<<<(unsigned)a.n, 32, smem, st>>>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap,
                                       a.result, (uint32_t)a.n);
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
int b200lz4_compress_algo = 2;        // 2 = decoupled lookup/parse (default), 1 = the original coupled warp parser
}

cudaError_t launch_compress_fast(const BatchArgs& a, int max_src_len, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    const bool u16 = max_src_len > 0 && max_src_len <= 65536;
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

} // namespace b200
