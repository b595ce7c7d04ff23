// lz4hc_compress.cu — placeholder until the HC kernel lands; keeps the C ABI complete.
#include "kernels.h"
namespace b200 {
cudaError_t launch_compress_hc(const BatchArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
}
