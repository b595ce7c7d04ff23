// misc_harness.cpp — the HC compressor and the long-stream / streaming XXH kernels on the SIMT emulator (simt.h),
// exported with a C ABI for tests/test_kernel_logic_cpu.py.  Test infrastructure only.
#include "../../lz4-java_b200/csrc/lz4hc_compress.cu"
#include "../../lz4-java_b200/csrc/xxhash.cu"

using namespace b200;

extern "C" int sim_compress_hc(const uint8_t* src, int n, uint8_t* dst, int cap, int level, int bucket_log, int ways)
{
    uint64_t zero = 0; int32_t sl = n, dc = cap, res = 0x7FFFFFFF;
#define ARGS src, &zero, &sl, dst, &zero, &dc, &res, 1u, level
    if (ways == 16) { if (bucket_log == 10) simt::launch(1, 128, [&] { lz4hc_compress_kernel<10, 16>(ARGS); }); else simt::launch(1, 128, [&] { lz4hc_compress_kernel<11, 16>(ARGS); }); }
    else            { if (bucket_log == 10) simt::launch(1, 128, [&] { lz4hc_compress_kernel<10, 32>(ARGS); }); else simt::launch(1, 128, [&] { lz4hc_compress_kernel<11, 32>(ARGS); }); }
#undef ARGS
    return res;
}

extern "C" uint32_t sim_xxh32_long(const uint8_t* p, int len, uint32_t seed)
{
    uint64_t zero = 0; int32_t l = len; uint32_t out = 0;
    simt::launch(1, 32, [&] { xxh32_long_kernel(p, &zero, &l, seed, &out, 1u); });
    return out;
}
extern "C" uint64_t sim_xxh64_long(const uint8_t* p, int len, uint64_t seed)
{
    uint64_t zero = 0; int32_t l = len; uint64_t out = 0;
    simt::launch(1, 32, [&] { xxh64_long_kernel(p, &zero, &l, seed, &out, 1u); });
    return out;
}
// streaming: reset, then one update per chunk boundary given in cuts[], then digest
extern "C" uint32_t sim_xxh32_stream(const uint8_t* p, int len, uint32_t seed, const int* cuts, int ncuts)
{
    Xxh32State st; std::memset(&st, 0, sizeof st);
    simt::launch(1, 32, [&] { xxh32_stream_kernel(&st, XXH_OP_RESET, seed, nullptr, 0); });
    int pos = 0;
    for (int k = 0; k <= ncuts; k++) {
        const int end = k < ncuts ? cuts[k] : len;
        simt::launch(1, 32, [&] { xxh32_stream_kernel(&st, XXH_OP_UPDATE, 0, p + pos, (size_t)(end - pos)); });
        pos = end;
        simt::launch(1, 32, [&] { xxh32_stream_kernel(&st, XXH_OP_DIGEST, 0, nullptr, 0); });      // digest is non-destructive
    }
    return st.digest;
}
extern "C" uint64_t sim_xxh64_stream(const uint8_t* p, int len, uint64_t seed, const int* cuts, int ncuts)
{
    Xxh64State st; std::memset(&st, 0, sizeof st);
    simt::launch(1, 32, [&] { xxh64_stream_kernel(&st, XXH_OP_RESET, seed, nullptr, 0); });
    int pos = 0;
    for (int k = 0; k <= ncuts; k++) {
        const int end = k < ncuts ? cuts[k] : len;
        simt::launch(1, 32, [&] { xxh64_stream_kernel(&st, XXH_OP_UPDATE, 0, p + pos, (size_t)(end - pos)); });
        pos = end;
        simt::launch(1, 32, [&] { xxh64_stream_kernel(&st, XXH_OP_DIGEST, 0, nullptr, 0); });
    }
    return st.digest;
}


// ---- compaction (scan + gather)
#include "../../lz4-java_b200/csrc/compact.cu"

extern "C" void sim_compact(const uint8_t* slots, const uint64_t* slot_off, const int32_t* lens, uint8_t* out, uint64_t* out_off, uint64_t* total, uint32_t n)
{
    simt::launch(1, 1024, [&] { compact_scan_kernel(lens, out_off, total, n); });
    simt::launch((n + 3) / 4, 128, [&] { compact_gather_kernel(slots, slot_off, lens, out, out_off, n); });
}
