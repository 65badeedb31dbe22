// Feature-cache index kernels (reference: cuda_cache.cu, dist_engine.cc:193-229).
//
// fgnn_k_cache_split is GetMissCacheIndex (count_miss_cache + 2x DeviceScan +
// get_miss_index + get_cache_index, cuda_cache.cu:33-234) as ONE chained
// single-pass launch: the node->slot table is read once per node instead of
// five times.
#include "common.cuh"

namespace fgnn {
namespace {

__global__ void __launch_bounds__(kBlock)
cache_table_fill_kernel(uint32_t *table, size_t num_nodes) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < num_nodes; i += stride)
    table[i] = kEmpty;
}

__global__ void __launch_bounds__(kBlock)
cache_table_scatter_kernel(uint32_t *table, const uint32_t *__restrict__ rank, size_t num_cached) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < num_cached; i += stride)
    table[rank[i]] = (uint32_t)i;
}

struct SplitSmem {
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

__global__ void __launch_bounds__(kBlock)
cache_split_kernel(const uint32_t *__restrict__ table, const uint32_t *__restrict__ nodes,
                   uint32_t n_max, const uint32_t *__restrict__ d_n,
                   uint32_t *__restrict__ miss_src, uint32_t *__restrict__ miss_dst,
                   uint32_t *__restrict__ cache_src, uint32_t *__restrict__ cache_dst,
                   uint32_t *d_counts, ChainWs *ws) {
  __shared__ SplitSmem sm;
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, kBlock, &begin, &end);

  unsigned long long partial = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock)
    partial += (__ldg(table + __ldg(nodes + i)) == kEmpty) ? 1u : 0u;
  unsigned long long chunk_miss;
  unsigned long long miss_base = chain_scan(ws, &sm.chain, p, partial, &chunk_miss);
  if (p == gridDim.x - 1 && threadIdx.x == 0) {
    d_counts[0] = (uint32_t)(miss_base + chunk_miss);
    d_counts[1] = n - (uint32_t)(miss_base + chunk_miss);
  }

  for (uint32_t t0 = begin; t0 < end; t0 += kBlock) {
    const uint32_t i = t0 + threadIdx.x;
    uint32_t node = 0, slot = 0, miss = 0;
    if (i < end) {
      node = __ldg(nodes + i);
      slot = __ldg(table + node);
      miss = slot == kEmpty ? 1u : 0u;
    }
    uint32_t tile_miss;
    const uint32_t excl = block_excl_scan(miss, sm.warp, &tile_miss);
    if (i < end) {
      if (miss) {  // cuda_cache.cu:96-102
        const uint32_t o = (uint32_t)miss_base + excl;
        miss_dst[o] = i;
        miss_src[o] = node;
      } else {     // cuda_cache.cu:140-146
        const uint32_t o = (t0 - (uint32_t)miss_base) + (threadIdx.x - excl);
        cache_dst[o] = i;
        cache_src[o] = slot;
      }
    }
    miss_base += tile_miss;
  }
  chain_finish(ws, &sm.chain);
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" int fgnn_k_cache_table_build(uint32_t *table, size_t num_nodes,
                                        const uint32_t *ranking_nodes, size_t num_cached,
                                        fgnn_stream_t stream) {
  if (!table || (num_cached > 0 && !ranking_nodes) || num_cached > num_nodes) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (num_nodes > 0) {
    cache_table_fill_kernel<<<persistent_grid(num_nodes, 4 * kBlock, 8, false), kBlock, 0, st>>>(
        table, num_nodes);
    note_launch();
  }
  if (num_cached > 0) {
    cache_table_scatter_kernel<<<persistent_grid(num_cached, 4 * kBlock, 8, false), kBlock, 0, st>>>(
        table, ranking_nodes, num_cached);
    note_launch();
  }
  return check_last();
}

extern "C" int fgnn_k_cache_split(const uint32_t *table, const uint32_t *nodes, uint32_t n_max,
                                  const uint32_t *d_n, uint32_t *miss_src, uint32_t *miss_dst,
                                  uint32_t *cache_src, uint32_t *cache_dst, uint32_t *d_counts,
                                  void *chain_ws, fgnn_stream_t stream) {
  if (!table || !d_counts || !chain_ws) return FGNN_ERR_BAD_ARG;
  if (n_max > 0 && (!nodes || !miss_src || !miss_dst || !cache_src || !cache_dst))
    return FGNN_ERR_BAD_ARG;
  const int grid = persistent_grid(n_max, kBlock, 8, true);
  cache_split_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(
      table, nodes, n_max, d_n, miss_src, miss_dst, cache_src, cache_dst, d_counts,
      (ChainWs *)chain_ws);
  note_launch();
  return check_last();
}
