// Library-level entry points of the kernel C-ABI.
#include "common.cuh"

namespace fgnn {
unsigned long long g_launch_count = 0;

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace fgnn

extern "C" const char *fgnn_k_version(void) { return "fgnn-b200 kernels r1 (sm_100a)"; }

extern "C" uint64_t fgnn_k_launch_count(void) { return fgnn::g_launch_count; }

extern "C" const char *fgnn_k_error_string(int code) {
  if (code == 0) return "ok";
  if (code == FGNN_ERR_BAD_ARG) return "fgnn: bad argument";
  if (code == FGNN_ERR_UNSUPPORTED) return "fgnn: unsupported configuration";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "fgnn: unknown error";
}
