// PreSC hotness counting and ranking (reference: cuda/pre_sampler.cc:39-142,
// dist/pre_sampler.cc:75-164).  The reference copies every batch's input_nodes
// to the host, counts with OpenMP and sorts 8-byte {freq,id} records with
// __gnu_parallel::sort; here counting is a fire-and-forget RED.ADD per node in
// HBM and the ranking is one descending radix sort of the same u64 records.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace fgnn {
namespace {

__global__ void __launch_bounds__(kBlock)
freq_count_kernel(uint32_t *freq, const uint32_t *__restrict__ nodes, uint32_t n_max,
                  const uint32_t *__restrict__ d_n) {
  const uint32_t n = load_count(n_max, d_n);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock)
    atomicAdd(freq + __ldg(nodes + i), 1u);
}

__global__ void __launch_bounds__(kBlock)
rank_keys_kernel(const uint32_t *__restrict__ freq, size_t num_nodes, unsigned long long *keys) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < num_nodes; i += stride)
    keys[i] = ((unsigned long long)freq[i] << 32) | (unsigned long long)i;  // pre_sampler.cc:44-49
}

__global__ void __launch_bounds__(kBlock)
rank_ids_kernel(const unsigned long long *__restrict__ keys, size_t num_nodes, uint32_t *rank) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < num_nodes; i += stride)
    rank[i] = (uint32_t)keys[i];  // pre_sampler.cc:121-131
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t sort_temp_bytes(size_t num_nodes) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned long long> keys(nullptr, nullptr);
  cub::DeviceRadixSort::SortKeysDescending(nullptr, bytes, keys, (int64_t)num_nodes, 0, 64,
                                           (cudaStream_t)0);
  return bytes;
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" int fgnn_k_freq_count(uint32_t *freq, const uint32_t *nodes, uint32_t n_max,
                                 const uint32_t *d_n, fgnn_stream_t stream) {
  if (n_max == 0) return 0;
  if (!freq || !nodes) return FGNN_ERR_BAD_ARG;
  freq_count_kernel<<<persistent_grid(n_max, kBlock, 8, false), kBlock, 0, (cudaStream_t)stream>>>(
      freq, nodes, n_max, d_n);
  note_launch();
  return check_last();
}

extern "C" size_t fgnn_k_presc_rank_workspace_bytes(size_t num_nodes) {
  return 2 * align256(num_nodes * sizeof(unsigned long long)) + align256(sort_temp_bytes(num_nodes));
}

extern "C" int fgnn_k_presc_rank(const uint32_t *freq, size_t num_nodes, uint32_t *ranking_nodes,
                                 void *workspace, size_t workspace_bytes, fgnn_stream_t stream) {
  if (num_nodes == 0) return 0;
  if (!freq || !ranking_nodes || !workspace) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_presc_rank_workspace_bytes(num_nodes)) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t kb = align256(num_nodes * sizeof(unsigned long long));
  unsigned long long *k0 = (unsigned long long *)workspace;
  unsigned long long *k1 = (unsigned long long *)((char *)workspace + kb);
  void *temp = (char *)workspace + 2 * kb;
  size_t temp_bytes = workspace_bytes - 2 * kb;
  const int grid = persistent_grid(num_nodes, 4 * kBlock, 8, false);
  rank_keys_kernel<<<grid, kBlock, 0, st>>>(freq, num_nodes, k0);
  cub::DoubleBuffer<unsigned long long> keys(k0, k1);
  cudaError_t e = cub::DeviceRadixSort::SortKeysDescending(temp, temp_bytes, keys,
                                                           (int64_t)num_nodes, 0, 64, st);
  if (e != cudaSuccess) return (int)e;
  rank_ids_kernel<<<grid, kBlock, 0, st>>>(keys.Current(), num_nodes, ranking_nodes);
  note_launch(3);
  return check_last();
}
