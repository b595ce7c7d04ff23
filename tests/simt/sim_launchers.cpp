// sim_launchers.cpp — the launch_* functions of kernels.h for the emulator build of the WHOLE library
// (tests/simt/build_sim_library.sh -> tests/simt/_build/libb200lz4_sim.so).  The host layer (capi.cu, frame.cu,
// containers.cu) is compiled unchanged against simt.h's stand-in CUDA runtime; the kernels are the product's own
// sources; only the <<<...>>> launch sites are restated here, with the same dispatch rules as the real launchers.
// TEST INFRASTRUCTURE ONLY: lets the host pipeline (chunking, bounce buffers, compaction, frame indexer, containers)
// and most of tests/test_gpu_parity.py run on a box without a GPU:
//     B200LZ4_TEST_SO=tests/simt/_build/libb200lz4_sim.so python -m pytest tests/test_gpu_parity.py -m gpu -k "not jni"
// The product never loads this library (it is not named libb200lz4.so and lives under tests/).
#include "../../lz4-java_b200/csrc/lz4_decompress.cu"
#include "../../lz4-java_b200/csrc/lz4_compress.cu"
#include "../../lz4-java_b200/csrc/lz4hc_compress.cu"
#include "../../lz4-java_b200/csrc/lz4hc2_compress.cu"
#include "../../lz4-java_b200/csrc/xxhash.cu"
#include "../../lz4-java_b200/csrc/compact.cu"

extern "C" {                       // the development knobs the real launchers define
int b200lz4_decompress_batch_below = 0x7FFFFFFF;
int b200lz4_compress_hash_log = 13, b200lz4_compress_stage = 0, b200lz4_compress_sparse = 0, b200lz4_compress_algo = 3, b200lz4_compress_subbatch = 16384;
int b200lz4_hc_bucket_log = 11, b200lz4_hc_ways = 32, b200lz4_hc_algo = 1, b200lz4_hc2_max_block = 262144, b200lz4_hc2_subbatch = 1184;
}

namespace b200 {

#define BA a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, (uint32_t)a.n

cudaError_t launch_decompress_safe(const BatchArgs& a, cudaStream_t)
{
    if (a.n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.n + 3) / 4);
    if (a.n < (size_t)b200lz4_decompress_batch_below) simt::launch(grid, 128, [&] { lz4_decompress_safe_kernel<4, true>(BA); });
    else simt::launch(grid, 128, [&] { lz4_decompress_safe_kernel<4, false>(BA); });
    return cudaSuccess;
}
cudaError_t launch_decompress_fast(const BatchArgs& a, cudaStream_t)
{
    if (a.n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.n + 3) / 4);
    if (a.n < (size_t)b200lz4_decompress_batch_below) simt::launch(grid, 128, [&] { lz4_decompress_fast_kernel<4, true>(BA); });
    else simt::launch(grid, 128, [&] { lz4_decompress_fast_kernel<4, false>(BA); });
    return cudaSuccess;
}

template <int HL, bool SPARSE> static void sim_v4(const BatchArgs& a)
{
    const uint32_t n = (uint32_t)a.n;
    std::vector<uint16_t> dist((size_t)n * K4_DIST_STRIDE); std::vector<uint32_t> mask((size_t)n * K4_MASK_STRIDE);
    std::vector<uint2> rec((size_t)n * K4_REC_STRIDE); std::vector<int32_t> cnt(n);
    simt::launch(n, 32, [&] { lz4c4_lookup_kernel<HL, SPARSE>(a.src_base, a.src_off, a.src_len, 0u, n, dist.data(), mask.data()); });
    simt::launch((n + 127) / 128, 128, [&] { lz4c4_walk_kernel(a.src_base, a.src_off, a.src_len, 0u, n, dist.data(), mask.data(), rec.data(), cnt.data()); });
    simt::launch((n + 3) / 4, 128, [&] { lz4c4_layout_kernel(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, 0u, n, rec.data(), cnt.data()); });
}
template <int HL, bool U16> static void sim_v2(const BatchArgs& a) { simt::launch((unsigned)a.n, 32, [&] { lz4_compress_fast2_kernel<HL, U16>(BA); }); }
template <int HL, bool SP> static void sim_v3(const BatchArgs& a) { simt::launch((unsigned)a.n, 32 * B200_V3_WARPS, [&] { lz4_compress_fast3_kernel<HL, SP>(BA); }); }
template <int HL, bool U16, bool ST> static void sim_v1(const BatchArgs& a) { simt::launch((unsigned)a.n, 32, [&] { lz4_compress_fast_kernel<HL, U16, ST>(BA); }); }

cudaError_t launch_compress_fast(const BatchArgs& a, int max_src_len, cudaStream_t)      // same dispatch as lz4_compress.cu
{
    if (a.n == 0) return cudaSuccess;
    const bool u16 = max_src_len > 0 && max_src_len <= 65536;
    const int algo = b200lz4_compress_algo, hl = b200lz4_compress_hash_log, sp = b200lz4_compress_sparse, stg = b200lz4_compress_stage;
    if (algo == 4 && !stg && u16) {
        if (hl == 12) { if (sp) sim_v4<12, true>(a); else sim_v4<12, false>(a); } else { if (sp) sim_v4<13, true>(a); else sim_v4<13, false>(a); }
    } else if ((algo == 3 || algo == 4) && !stg) {
        if (!u16) sim_v2<12, false>(a);
        else if (hl == 12) { if (sp) sim_v3<12, true>(a); else sim_v3<12, false>(a); }
        else { if (sp) sim_v3<13, true>(a); else sim_v3<13, false>(a); }
    } else if (algo == 2 && !stg) {
        if (!u16) sim_v2<12, false>(a); else if (hl == 12) sim_v2<12, true>(a); else sim_v2<13, true>(a);
    } else if (u16) {
        if (stg) { if (hl == 12) sim_v1<12, true, true>(a); else sim_v1<13, true, true>(a); }
        else if (hl == 12) sim_v1<12, true, false>(a); else if (hl == 11) sim_v1<11, true, false>(a); else sim_v1<13, true, false>(a);
    } else sim_v1<12, false, false>(a);
    return cudaSuccess;
}

cudaError_t launch_compress_hc2(const BatchArgs& a, cudaStream_t)
{
    const uint32_t n = (uint32_t)a.n, stride = (uint32_t)((b200lz4_hc2_max_block + 3) & ~3), rec_stride = stride / 4 + 16;
    for (uint32_t first = 0; first < n; first++) {                      // one block per round keeps the scratch small
        std::vector<uint32_t> best(stride), cost(stride); std::vector<Hc2Rec> rec(rec_stride); int32_t cnt = 0;
        simt::launch(1, HC2_THREADS, [&] { lz4hc2_search_kernel(a.src_base, a.src_off, a.src_len, first, 1u, best.data(), stride); });
        simt::launch(1, 128, [&] { lz4hc2_parse_kernel(a.src_base, a.src_off, a.src_len, first, 1u, best.data(), cost.data(), stride, rec.data(), rec_stride, &cnt); });
        simt::launch(1, 128, [&] { lz4hc2_layout_kernel(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, first, 1u, rec.data(), rec_stride, &cnt); });
    }
    return cudaSuccess;
}
cudaError_t launch_compress_hc(const BatchArgs& a, int level, cudaStream_t st)
{
    if (a.n == 0) return cudaSuccess;
    if (level < 1) level = 9;
    if (b200lz4_hc_algo == 2) return launch_compress_hc2(a, st);
    const unsigned n = (unsigned)a.n;
#define HCA a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, (uint32_t)a.n, level
    if (b200lz4_hc_ways == 16) { if (b200lz4_hc_bucket_log == 10) simt::launch(n, 128, [&] { lz4hc_compress_kernel<10, 16>(HCA); }); else simt::launch(n, 128, [&] { lz4hc_compress_kernel<11, 16>(HCA); }); }
    else                       { if (b200lz4_hc_bucket_log == 10) simt::launch(n, 128, [&] { lz4hc_compress_kernel<10, 32>(HCA); }); else simt::launch(n, 128, [&] { lz4hc_compress_kernel<11, 32>(HCA); }); }
#undef HCA
    return cudaSuccess;
}

// The TMA-staged batch kernel is PTX; the emulator build hashes every buffer with the one-warp-per-stream kernels.
cudaError_t launch_xxh32_long(const uint8_t* base, const uint64_t* off, const int32_t* len, uint32_t seed, uint32_t* out, size_t n, cudaStream_t)
{ if (n) simt::launch((unsigned)n, 32, [&] { xxh32_long_kernel(base, off, len, seed, out, (uint32_t)n); }); return cudaSuccess; }
cudaError_t launch_xxh64_long(const uint8_t* base, const uint64_t* off, const int32_t* len, uint64_t seed, uint64_t* out, size_t n, cudaStream_t)
{ if (n) simt::launch((unsigned)n, 32, [&] { xxh64_long_kernel(base, off, len, seed, out, (uint32_t)n); }); return cudaSuccess; }
cudaError_t launch_xxh32(const uint8_t* base, const uint64_t* off, const int32_t* len, uint32_t seed, uint32_t* out, size_t n, cudaStream_t st)
{ return launch_xxh32_long(base, off, len, seed, out, n, st); }
cudaError_t launch_xxh64(const uint8_t* base, const uint64_t* off, const int32_t* len, uint64_t seed, uint64_t* out, size_t n, cudaStream_t st)
{ return launch_xxh64_long(base, off, len, seed, out, n, st); }
cudaError_t launch_xxh32_stream(Xxh32State* s, int op, uint32_t seed, const uint8_t* data, size_t len, cudaStream_t)
{ simt::launch(1, 32, [&] { xxh32_stream_kernel(s, op, seed, data, len); }); return cudaSuccess; }
cudaError_t launch_xxh64_stream(Xxh64State* s, int op, uint64_t seed, const uint8_t* data, size_t len, cudaStream_t)
{ simt::launch(1, 32, [&] { xxh64_stream_kernel(s, op, seed, data, len); }); return cudaSuccess; }

cudaError_t launch_compact(const uint8_t* slots, const uint64_t* slot_off, const int32_t* lens, uint8_t* out, uint64_t* out_off, uint64_t* total, size_t n, cudaStream_t)
{
    if (n == 0) return cudaSuccess;
    simt::launch(1, 1024, [&] { compact_scan_kernel(lens, out_off, total, (uint32_t)n); });
    simt::launch((unsigned)((n + 3) / 4), 128, [&] { compact_gather_kernel(slots, slot_off, lens, out, out_off, (uint32_t)n); });
    return cudaSuccess;
}
cudaError_t launch_gather(const uint8_t* src, const uint64_t* src_off, const int32_t* lens, uint8_t* dst, const uint64_t* dst_off, size_t n, cudaStream_t)
{
    if (n) simt::launch((unsigned)((n + 3) / 4), 128, [&] { compact_gather_kernel(src, src_off, lens, dst, dst_off, (uint32_t)n); });
    return cudaSuccess;
}

} // namespace b200
