// comp_harness.cpp — the fast-compress kernels of lz4-java_b200/csrc/lz4_compress.cu on the SIMT emulator (simt.h),
// exported with a C ABI for tests/test_kernel_logic_cpu.py.  Test infrastructure only.
#include "../../lz4-java_b200/csrc/lz4_compress.cu"

using namespace b200;

// algo: 3 = two-warp pipeline (the default), 2 = one-warp decoupled, 1 = coupled parser (stage: shared-memory input)
extern "C" int sim_compress_fast(const uint8_t* src, int n, uint8_t* dst, int cap, int algo, int hash_log, int u16, int sparse, int stage)
{
    uint64_t zero = 0; int32_t sl = n, dc = cap, res = 0x7FFFFFFF;
#define ARGS src, &zero, &sl, dst, &zero, &dc, &res, 1u
    if (algo == 5) {          // stage = 100 * warps + 20 + chunk buffers
        if (stage == 322) simt::launch(1, 96, [&] { lz4_compress_wide_kernel<13, 2, 2, 3, 1>(ARGS); });
        else simt::launch(1, 64, [&] { lz4_compress_wide_kernel<13, 2, 2, 2, 1>(ARGS); });
    } else if (algo == 3) {
        if (hash_log == 12) { if (sparse) simt::launch(1, 64, [&] { lz4_compress_fast3_kernel<12, true>(ARGS); }); else simt::launch(1, 64, [&] { lz4_compress_fast3_kernel<12, false>(ARGS); }); }
        else                { if (sparse) simt::launch(1, 64, [&] { lz4_compress_fast3_kernel<13, true>(ARGS); }); else simt::launch(1, 64, [&] { lz4_compress_fast3_kernel<13, false>(ARGS); }); }
    } else if (algo == 2) {
        if (!u16) simt::launch(1, 32, [&] { lz4_compress_fast2_kernel<12, false>(ARGS); });
        else if (hash_log == 12) simt::launch(1, 32, [&] { lz4_compress_fast2_kernel<12, true>(ARGS); });
        else simt::launch(1, 32, [&] { lz4_compress_fast2_kernel<13, true>(ARGS); });
    } else {
        if (!u16) simt::launch(1, 32, [&] { lz4_compress_fast_kernel<12, false, false>(ARGS); });
        else if (stage) simt::launch(1, 32, [&] { lz4_compress_fast_kernel<13, true, true>(ARGS); });
        else if (hash_log == 12) simt::launch(1, 32, [&] { lz4_compress_fast_kernel<12, true, false>(ARGS); });
        else simt::launch(1, 32, [&] { lz4_compress_fast_kernel<13, true, false>(ARGS); });
    }
#undef ARGS
    return res;
}
