// comp_harness.cpp — the fast-compress kernels of lz4-java_b200/csrc on the SIMT emulator (simt.h), exported with a C ABI
// for tests/test_kernel_logic_cpu.py.  Test infrastructure only.
#include "../../lz4-java_b200/csrc/lz4_compress.cu"

using namespace b200;

// kind: 3 / 2 = the <= 64 KiB kernel with three / two warps per block; 0 = the long-block kernel (32-bit table, any size)
extern "C" int sim_compress_fast(const uint8_t* src, int n, uint8_t* dst, int cap, int kind)
{
    uint64_t zero = 0; int32_t sl = n, dc = cap, res = 0x7FFFFFFF;
#define ARGS src, &zero, &sl, dst, &zero, &dc, &res, 1u
    if (kind == 3) simt::launch(1, 96, [&] { lz4_compress_wide_kernel<13, 2, 2, 3, 1>(ARGS); });
    else if (kind == 2) simt::launch(1, 64, [&] { lz4_compress_wide_kernel<13, 2, 2, 2, 1>(ARGS); });
    else simt::launch(1, 32, [&] { lz4_compress_long_kernel<12, false>(ARGS); });
#undef ARGS
    return res;
}
