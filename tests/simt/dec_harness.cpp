// dec_harness.cpp — the decoder kernels of lz4-java_b200/csrc/lz4_decompress.cu compiled as host C++ on the
// SIMT emulator (simt.h), exported with a C ABI for tests/test_kernel_logic_cpu.py.  Test infrastructure only.
//   g++ -O1 -std=c++17 -shared -fPIC -DB200_HOST_SIM -Itests/simt -Ilz4-java_b200/csrc tests/simt/dec_harness.cpp
#include "../../lz4-java_b200/csrc/lz4_decompress.cu"

using namespace b200;

template <bool BATCH>
static int run_safe(const uint8_t* src, int n, uint8_t* dst, int cap)
{
    uint64_t zero = 0; int32_t sl = n, dc = cap, res = 0x7FFFFFFF;
    simt::launch(1, 32, [&] { lz4_decompress_safe_kernel<1, BATCH>(src, &zero, &sl, dst, &zero, &dc, &res, 1u); });
    return res;
}
template <bool BATCH>
static int run_fast(const uint8_t* src, int avail, uint8_t* dst, int n)
{
    uint64_t zero = 0; int32_t av = avail, dl = n, res = 0x7FFFFFFF;
    simt::launch(1, 32, [&] { lz4_decompress_fast_kernel<1, BATCH>(src, &zero, &av, dst, &zero, &dl, &res, 1u); });
    return res;
}

extern "C" {
int sim_decompress_safe(const uint8_t* src, int n, uint8_t* dst, int cap, int batched)
{ return batched ? run_safe<true>(src, n, dst, cap) : run_safe<false>(src, n, dst, cap); }
int sim_decompress_fast(const uint8_t* src, int avail, uint8_t* dst, int n, int batched)
{ return batched ? run_fast<true>(src, avail, dst, n) : run_fast<false>(src, avail, dst, n); }
}
