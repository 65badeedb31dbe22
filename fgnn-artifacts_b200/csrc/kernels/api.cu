// Library-level entry points of the kernel C-ABI.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace fgnn {
unsigned long long g_launch_count = 0;

// ---- launch event trace (see common.cuh) ----
bool g_trace_on = false;
namespace {
struct TraceRec {
  cudaEvent_t ev;
  int label;
  cudaStream_t stream;
};
std::mutex g_trace_mu;
std::vector<TraceRec> g_trace;
cudaEvent_t g_trace_base = nullptr;
size_t g_trace_cap = 0;
}  // namespace
void trace_mark_slow(cudaStream_t st, int label) {
  std::lock_guard<std::mutex> lk(g_trace_mu);
  if (!g_trace_on || g_trace.size() >= g_trace_cap) return;
  TraceRec r{nullptr, label, st};
  if (cudaEventCreate(&r.ev) != cudaSuccess) return;
  if (cudaEventRecord(r.ev, st) != cudaSuccess) {
    cudaEventDestroy(r.ev);
    return;
  }
  g_trace.push_back(r);
}

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
int grid_share_div() {
  static int d = 0;
  const char *dyn = getenv("FGNN_TUNING_DYNAMIC");  // sweeps: re-read per call
  if (d == 0 || (dyn && atoi(dyn) != 0)) {
    const char *v = getenv("FGNN_GRID_DIV");
    d = v && *v ? atoi(v) : 1;
    if (d < 1) d = 1;
  }
  return d;
}
}  // namespace fgnn

extern "C" int fgnn_k_trace_enable(int on, size_t max_records) {
  std::lock_guard<std::mutex> lk(fgnn::g_trace_mu);
  for (auto &r : fgnn::g_trace) cudaEventDestroy(r.ev);
  fgnn::g_trace.clear();
  if (fgnn::g_trace_base) {
    cudaEventDestroy(fgnn::g_trace_base);
    fgnn::g_trace_base = nullptr;
  }
  fgnn::g_trace_on = false;
  if (!on) return 0;
  cudaError_t e = cudaEventCreate(&fgnn::g_trace_base);
  if (e != cudaSuccess) return (int)e;
  e = cudaEventRecord(fgnn::g_trace_base, (cudaStream_t)0);  // time zero: everything enqueued so far has to drain first
  if (e != cudaSuccess) return (int)e;
  fgnn::g_trace_cap = max_records ? max_records : 4096;
  fgnn::g_trace.reserve(fgnn::g_trace_cap);
  fgnn::g_trace_on = true;
  return 0;
}

extern "C" long fgnn_k_trace_dump(size_t max_records, int *labels, uint64_t *streams, float *ms) {
  std::lock_guard<std::mutex> lk(fgnn::g_trace_mu);
  fgnn::g_trace_on = false;
  if (!fgnn::g_trace_base) return 0;
  size_t n = 0;
  for (auto &r : fgnn::g_trace) {
    float t = -1.0f;
    if (cudaEventSynchronize(r.ev) == cudaSuccess) cudaEventElapsedTime(&t, fgnn::g_trace_base, r.ev);
    if (n < max_records && labels && streams && ms) {
      labels[n] = r.label;
      streams[n] = (uint64_t)(uintptr_t)r.stream;
      ms[n] = t;
      ++n;
    }
    cudaEventDestroy(r.ev);
  }
  fgnn::g_trace.clear();
  cudaEventDestroy(fgnn::g_trace_base);
  fgnn::g_trace_base = nullptr;
  cudaGetLastError();
  return (long)n;
}

extern "C" const char *fgnn_k_version(void) { return "fgnn-b200 kernels r1 (sm_100a)"; }

extern "C" uint64_t fgnn_k_launch_count(void) { return fgnn::g_launch_count; }

extern "C" const char *fgnn_k_error_string(int code) {
  if (code == 0) return "ok";
  if (code == FGNN_ERR_BAD_ARG) return "fgnn: bad argument";
  if (code == FGNN_ERR_UNSUPPORTED) return "fgnn: unsupported configuration";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "fgnn: unknown error";
}

// ---- device-memory plumbing for partitioned cache shards ----------------------------------
extern "C" int fgnn_k_shard_alloc(void **ptr, size_t bytes) {
  if (!ptr) return FGNN_ERR_BAD_ARG;
  return (int)cudaMalloc(ptr, bytes ? bytes : 256);  // plain cudaMalloc: exportable through CUDA IPC
}
extern "C" int fgnn_k_shard_free(void *ptr) { return (int)cudaFree(ptr); }
extern "C" int fgnn_k_ipc_export(void *ptr, void *handle64) {
  if (!ptr || !handle64) return FGNN_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == FGNN_IPC_HANDLE_BYTES, "ipc handle size");
  return (int)cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle64, ptr);
}
extern "C" int fgnn_k_ipc_open(const void *handle64, void **ptr) {
  if (!ptr || !handle64) return FGNN_ERR_BAD_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int fgnn_k_ipc_close(void *ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
extern "C" int fgnn_k_enable_peer(int peer_device) {
  int cur = 0;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return (int)e;
  if (cur == peer_device) return 0;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (e != cudaSuccess) return (int)e;
  if (!can) return FGNN_ERR_UNSUPPORTED;
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return 0;
  }
  return (int)e;
}
