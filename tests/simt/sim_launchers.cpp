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
#include "../../lz4-java_b200/csrc/xxhash.cu"
#include "../../lz4-java_b200/csrc/compact.cu"

namespace b200 {

#define BA a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, (uint32_t)a.n

cudaError_t launch_decompress_safe(const BatchArgs& a, cudaStream_t)
{
    if (a.n) simt::launch((unsigned)((a.n + 3) / 4), 128, [&] { lz4_decompress_safe_kernel<4, true>(BA); });
    return cudaSuccess;
}
cudaError_t launch_decompress_fast(const BatchArgs& a, cudaStream_t)
{
    if (a.n) simt::launch((unsigned)((a.n + 3) / 4), 128, [&] { lz4_decompress_fast_kernel<4, true>(BA); });
    return cudaSuccess;
}

cudaError_t launch_compress_fast(const BatchArgs& a, int max_src_len, cudaStream_t)      // same dispatch as lz4_compress.cu
{
    if (a.n == 0) return cudaSuccess;
    if (max_src_len > 0 && max_src_len <= 65536) simt::launch((unsigned)a.n, 96, [&] { lz4_compress_wide_kernel<13, 2, 2, 3, 1>(BA); });
    else simt::launch((unsigned)a.n, 32, [&] { lz4_compress_long_kernel<12, false>(BA); });
    return cudaSuccess;
}

cudaError_t launch_compress_hc(const BatchArgs& a, int level, cudaStream_t)
{
    if (a.n == 0) return cudaSuccess;
    if (level < 1) level = 9;
    simt::launch((unsigned)a.n, 128, [&] { lz4hc_compress_kernel<11, 32>(a.src_base, a.src_off, a.src_len, a.dst_base, a.dst_off, a.dst_cap, a.result, (uint32_t)a.n, level); });
    return cudaSuccess;
}

// The TMA-staged batch kernel is PTX; the emulator build hashes every buffer with the one-warp-per-stream kernels.
cudaError_t launch_xxh32_long(const uint8_t* base, const uint64_t* off, const int32_t* len, uint32_t seed, uint32_t* out, size_t n, cudaStream_t)
{ if (n) simt::launch((unsigned)n, 32, [&] { xxh32_long_kernel(base, off, len, seed, out, (uint32_t)n); }); return cudaSuccess; }
cudaError_t launch_xxh64_long(const uint8_t* base, const uint64_t* off, const int32_t* len, uint64_t seed, uint64_t* out, size_t n, cudaStream_t)
{ if (n) simt::launch((unsigned)n, 32, [&] { xxh64_long_kernel(base, off, len, seed, out, (uint32_t)n); }); return cudaSuccess; }
cudaError_t launch_xxh32(const uint8_t* base, const uint64_t* off, const int32_t* len, uint32_t seed, uint32_t* out, size_t n, cudaStream_t st)
{ return launch_xxh32_long(base, off, len, seed, out, n, st); }
cudaError_t launch_xxh32_frames_chained(const uint8_t* slots, const uint64_t* blk_off, const uint32_t* f_first, const uint32_t* f_nblk,
                                        const int32_t* blk_comp, const int32_t* blk_rawlen, const int32_t* c_res, uint32_t* out, size_t n, cudaStream_t)
{   // (the emulator runs launches one after the other: the decoder has finished, nothing spins)
    if (n) simt::launch((unsigned)n, 32, [&] { xxh32_frames_chained_kernel(slots, blk_off, f_first, f_nblk, blk_comp, blk_rawlen, c_res, out, (uint32_t)n); });
    return cudaSuccess;
}
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
