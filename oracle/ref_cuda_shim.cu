// TEST INFRASTRUCTURE — not product code.
//
// Lets the reference's *unmodified* CUDA hot-path translation units
//   samgraph/common/cuda/cuda_sampling_{khop0,khop1,khop2,weighted_khop,weighted_khop_prefix,
//   weighted_khop_hash_dedup,random_walk}.cu, cuda_frequency_hashmap.cu, cuda_hashtable.cu, cuda_mapping.cu,
//   cuda_cache.cu, cuda_random_states.cu
// (compiled IN PLACE from /root/reference for sm_100a by `make -C oracle refcuda`, never copied) link and run
// without the reference's engine: it supplies
//   * samgraph::common::Device over plain cudaMalloc / malloc (interface: device.h:36-62; the reference's
//     cuda_device.cc + workspace_pool.cc pull in the whole engine),
//   * the two Profiler entry points those files call (profiler.cc depends on Engine and PreSampler),
//   * a flat extern "C" surface taking raw DEVICE pointers, so the GPU tests can feed identical inputs to the
//     reference kernels (cuRAND XORWOW, wall-clock seeded: compared distributionally) and to ours.
// Output: oracle/_ref/libsamgraph_ref_cuda.so.  Used ONLY by tests/ (-m gpu) and tools/ref_cuda_timing.py.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "samgraph/common/common.h"
#include "samgraph/common/constant.h"
#include "samgraph/common/cuda/cuda_frequency_hashmap.h"
#include "samgraph/common/cuda/cuda_function.h"
#include "samgraph/common/cuda/cuda_hashtable.h"
#include "samgraph/common/cuda/cuda_random_states.h"
#include "samgraph/common/device.h"
#include "samgraph/common/profiler.h"
#include "samgraph/common/run_config.h"

#define SHIM_CUDA(x)                                                                          \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      fprintf(stderr, "ref_cuda_shim: %s failed: %s\n", #x, cudaGetErrorString(e_));           \
      abort();                                                                                \
    }                                                                                         \
  } while (0)

namespace samgraph {
namespace common {

namespace {
class ShimGpuDevice final : public Device {
 public:
  void SetDevice(Context ctx) override { SHIM_CUDA(cudaSetDevice(ctx.device_id)); }
  void *AllocDataSpace(Context ctx, size_t nbytes, size_t) override {
    void *p = nullptr;
    SHIM_CUDA(cudaSetDevice(ctx.device_id));
    SHIM_CUDA(cudaMalloc(&p, nbytes ? nbytes : 64));
    return p;
  }
  void FreeDataSpace(Context ctx, void *ptr) override {
    SHIM_CUDA(cudaSetDevice(ctx.device_id));
    SHIM_CUDA(cudaFree(ptr));
  }
  void CopyDataFromTo(const void *from, size_t from_offset, void *to, size_t to_offset, size_t nbytes, Context,
                      Context, StreamHandle stream) override {
    SHIM_CUDA(cudaMemcpyAsync(static_cast<char *>(to) + to_offset, static_cast<const char *>(from) + from_offset,
                              nbytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  }
  void StreamSync(Context, StreamHandle stream) override {
    SHIM_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  }
};
class ShimCpuDevice final : public Device {
 public:
  void SetDevice(Context) override {}
  void *AllocDataSpace(Context, size_t nbytes, size_t alignment) override {
    void *p = nullptr;
    if (posix_memalign(&p, alignment < 64 ? 64 : alignment, nbytes ? nbytes : 64) != 0) abort();
    return p;
  }
  void FreeDataSpace(Context, void *ptr) override { free(ptr); }
  void CopyDataFromTo(const void *from, size_t fo, void *to, size_t to_off, size_t nbytes, Context, Context,
                      StreamHandle) override {
    memcpy(static_cast<char *>(to) + to_off, static_cast<const char *>(from) + fo, nbytes);
  }
  void StreamSync(Context, StreamHandle) override {}
};
}  // namespace

void *Device::AllocWorkspace(Context ctx, size_t nbytes, double) { return AllocDataSpace(ctx, nbytes, kTempAllocaAlignment); }
void Device::FreeWorkspace(Context ctx, void *ptr, size_t) { FreeDataSpace(ctx, ptr); }
StreamHandle Device::CreateStream(Context) { return nullptr; }
void Device::FreeStream(Context, StreamHandle) {}
void Device::SyncStreamFromTo(Context, StreamHandle, StreamHandle) {}
Device *Device::Get(Context ctx) {
  static ShimGpuDevice gpu;
  static ShimCpuDevice cpu;
  return ctx.device_type == kGPU ? static_cast<Device *>(&gpu) : static_cast<Device *>(&cpu);
}

// profiler.cc needs Engine and PreSampler; the kernels' host wrappers only add step timings
TraceEvent::TraceEvent() : begin(0), end(0) {}
TraceData::TraceData(size_t n) : events(n) {}
Profiler::Profiler() {}
Profiler &Profiler::Get() {
  static Profiler p;
  return p;
}
void Profiler::LogStepAdd(uint64_t, LogStepItem, double) {}

}  // namespace common
}  // namespace samgraph

namespace sc = samgraph::common;
using sc::IdType;

namespace {
int g_dev = 0;
sc::Context Gpu() { return sc::Context(sc::kGPU, g_dev); }
size_t ReadCount(size_t *d) {
  size_t h = 0;
  SHIM_CUDA(cudaMemcpy(&h, d, sizeof(size_t), cudaMemcpyDeviceToHost));
  return h;
}
}  // namespace

extern "C" {

void refcuda_set_device(int dev) {
  g_dev = dev;
  SHIM_CUDA(cudaSetDevice(dev));
}

// GPURandomStates (cuda_random_states.cu:64-108): XORWOW states seeded from the wall clock
void *refcuda_states_new(int sample_type, const size_t *fanout, size_t num_fanout, size_t batch_size,
                         size_t num_random_walk) {
  sc::RunConfig::num_random_walk = num_random_walk;
  std::vector<size_t> f(fanout, fanout + num_fanout);
  auto *s = new sc::cuda::GPURandomStates(static_cast<sc::SampleType>(sample_type), f, batch_size, Gpu());
  SHIM_CUDA(cudaDeviceSynchronize());
  return s;
}
void refcuda_states_free(void *s) { delete static_cast<sc::cuda::GPURandomStates *>(s); }

// kind = SampleType (common.h:50-58): 0 khop0, 1 khop1, 2 weighted alias, 4 weighted prefix, 5 khop2,
// 6 weighted hash-dedup.  All pointers are device pointers; returns the number of sampled edges.
// khop2 mutates `indices`, hash-dedup takes non-const tables: the caller passes scratch copies.
size_t refcuda_sample(int kind, const IdType *indptr, IdType *indices, float *prob, const IdType *alias,
                      const float *prefix, const IdType *input, size_t num_input, size_t fanout, IdType *out_src,
                      IdType *out_dst, void *states) {
  auto *rs = static_cast<sc::cuda::GPURandomStates *>(states);
  size_t *num_out = nullptr;
  SHIM_CUDA(cudaMalloc(&num_out, sizeof(size_t)));
  SHIM_CUDA(cudaMemset(num_out, 0, sizeof(size_t)));
  const sc::Context ctx = Gpu();
  switch (kind) {
    case 0: sc::cuda::GPUSampleKHop0(indptr, indices, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    case 1: sc::cuda::GPUSampleKHop1(indptr, indices, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    case 2: sc::cuda::GPUSampleWeightedKHop(indptr, indices, prob, alias, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    case 4: sc::cuda::GPUSampleWeightedKHopPrefix(indptr, indices, prefix, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    case 5: sc::cuda::GPUSampleKHop2(indptr, indices, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    case 6: sc::cuda::GPUSampleWeightedKHopHashDedup(indptr, indices, prob, alias, input, num_input, fanout, out_src, out_dst, num_out, ctx, nullptr, rs, 0); break;
    default: fprintf(stderr, "refcuda_sample: bad kind %d\n", kind); abort();
  }
  SHIM_CUDA(cudaDeviceSynchronize());
  const size_t n = ReadCount(num_out);
  SHIM_CUDA(cudaFree(num_out));
  return n;
}

// FrequencyHashmap (cuda_frequency_hashmap.cu) + GPUSampleRandomWalk (cuda_sampling_random_walk.cu:113-161)
void *refcuda_freqmap_new(size_t max_nodes, size_t edges_per_node) {
  return new sc::cuda::FrequencyHashmap(max_nodes, edges_per_node, Gpu());
}
void refcuda_freqmap_free(void *m) { delete static_cast<sc::cuda::FrequencyHashmap *>(m); }
size_t refcuda_random_walk(const IdType *indptr, const IdType *indices, const IdType *input, size_t num_input,
                           size_t walk_len, double restart_prob, size_t num_walk, size_t K, IdType *out_src,
                           IdType *out_dst, IdType *out_data, void *freqmap, void *states) {
  size_t *num_out = nullptr;
  SHIM_CUDA(cudaMalloc(&num_out, sizeof(size_t)));
  SHIM_CUDA(cudaMemset(num_out, 0, sizeof(size_t)));
  sc::cuda::GPUSampleRandomWalk(indptr, indices, input, num_input, walk_len, restart_prob, num_walk, K, out_src,
                                out_dst, out_data, num_out, static_cast<sc::cuda::FrequencyHashmap *>(freqmap),
                                Gpu(), nullptr, static_cast<sc::cuda::GPURandomStates *>(states), 0);
  SHIM_CUDA(cudaDeviceSynchronize());
  const size_t n = ReadCount(num_out);
  SHIM_CUDA(cudaFree(num_out));
  return n;
}
// GetTopK alone (cuda_frequency_hashmap.cu:1143-1367) on caller-provided walk edges: (src = start node's index
// in input_nodes... as produced by sample_random_walk, dst = visited node; EMPTY = walk ended).  The stock
// (SXN_REVISED) GetTopK overwrites both input arrays: the caller passes scratch copies.
size_t refcuda_topk(void *freqmap, IdType *in_src, IdType *in_dst, size_t num_edge,
                    const IdType *input_nodes, size_t num_nodes, size_t K, IdType *out_src, IdType *out_dst,
                    IdType *out_data) {
  size_t *num_out = nullptr;
  SHIM_CUDA(cudaMalloc(&num_out, sizeof(size_t)));
  SHIM_CUDA(cudaMemset(num_out, 0, sizeof(size_t)));
  static_cast<sc::cuda::FrequencyHashmap *>(freqmap)->GetTopK(in_src, in_dst, num_edge, input_nodes, num_nodes, K,
                                                             out_src, out_dst, out_data, num_out, nullptr, 0);
  SHIM_CUDA(cudaDeviceSynchronize());
  const size_t n = ReadCount(num_out);
  SHIM_CUDA(cudaFree(num_out));
  return n;
}

// OrderedHashTable (cuda_hashtable.cu) + GPUMapEdges (cuda_mapping.cu:68-81)
void *refcuda_ht_new(size_t size) { return new sc::cuda::OrderedHashTable(size, Gpu()); }
void refcuda_ht_free(void *t) { delete static_cast<sc::cuda::OrderedHashTable *>(t); }
void refcuda_ht_reset(void *t) {
  static_cast<sc::cuda::OrderedHashTable *>(t)->Reset(nullptr);
  SHIM_CUDA(cudaDeviceSynchronize());
}
void refcuda_ht_fill_unique(void *t, const IdType *input, size_t n) {
  static_cast<sc::cuda::OrderedHashTable *>(t)->FillWithUnique(input, n, nullptr);
  SHIM_CUDA(cudaDeviceSynchronize());
}
size_t refcuda_ht_fill_duplicates(void *t, const IdType *input, size_t n, IdType *unique) {
  IdType num_unique = 0;
  static_cast<sc::cuda::OrderedHashTable *>(t)->FillWithDuplicates(input, n, unique, &num_unique, nullptr);
  SHIM_CUDA(cudaDeviceSynchronize());
  return num_unique;
}
size_t refcuda_ht_num_items(void *t) { return static_cast<sc::cuda::OrderedHashTable *>(t)->NumItems(); }
void refcuda_map_edges(void *t, const IdType *src, IdType *new_src, const IdType *dst, IdType *new_dst, size_t n) {
  sc::cuda::GPUMapEdges(src, new_src, dst, new_dst, n, static_cast<sc::cuda::OrderedHashTable *>(t)->DeviceHandle(),
                        Gpu(), nullptr);
  SHIM_CUDA(cudaDeviceSynchronize());
}

// GetMissCacheIndex (cuda_cache.cu:162-234)
void refcuda_get_miss_cache_index(IdType *table, IdType *miss_src, IdType *miss_dst, size_t *num_miss,
                                  IdType *cache_src, IdType *cache_dst, size_t *num_cache, const IdType *nodes,
                                  size_t num_nodes) {
  sc::cuda::GetMissCacheIndex(table, Gpu(), miss_src, miss_dst, num_miss, cache_src, cache_dst, num_cache, nodes,
                              num_nodes, nullptr);
  SHIM_CUDA(cudaDeviceSynchronize());
}

}  // extern "C"
