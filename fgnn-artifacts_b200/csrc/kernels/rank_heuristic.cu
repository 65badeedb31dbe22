// cache_by_heuristic ranking on the GPU (reference: the offline CPU tool
// utility/data-process/toolkit/cache/cache_by_heuristic.cc:28-91, whose cache_by_heuristic.bin the engine
// loads at engine.cc:216-256):
//   1. the training nodes, in train_set order
//   2. their first-hop neighbours, in order of first appearance (train_set order, CSR order inside a row)
//   3. every other vertex by {out_degree, id} descending
// Steps 1+2 are exactly the ordered unique list of the sampler's hash table (FillWithUnique of the seeds, then
// FillWithDuplicates of the expanded neighbour lists), step 3 is the cache_by_degree ranking passed through the
// stable two-way split of GetMissCacheIndex with "already ranked" in the role of "cached".  So the only new
// kernels are the row expansion and two glue passes; everything runs on one stream with no host round trip.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace fgnn {
namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__global__ void __launch_bounds__(kBlock)
row_len_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ nodes, size_t n,
               uint32_t *__restrict__ lens) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    const uint32_t v = __ldg(nodes + i);
    lens[i] = __ldg(indptr + v + 1) - __ldg(indptr + v);
  }
}

__global__ void __launch_bounds__(kBlock)
row_len_sum_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ nodes, size_t n,
                   unsigned long long *d_total) {
  unsigned long long part = 0;
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) {
    const uint32_t v = __ldg(nodes + i);
    part += __ldg(indptr + v + 1) - __ldg(indptr + v);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, d);
  if ((threadIdx.x & 31) == 0 && part) atomicAdd(d_total, part);
}

// one warp per listed row: out[offs[i] + k] = indices[indptr[nodes[i]] + k]
__global__ void __launch_bounds__(kBlock)
expand_rows_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                   const uint32_t *__restrict__ nodes, size_t n, const uint32_t *__restrict__ offs,
                   uint32_t *__restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (kBlock / 32);
  for (size_t i = (size_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); i < n; i += warps) {
    const uint32_t v = __ldg(nodes + i);
    const uint32_t o = __ldg(indptr + v), len = __ldg(indptr + v + 1) - o;
    const uint32_t *src = indices + (size_t)o;
    uint32_t *dst = out + (size_t)__ldg(offs + i);
    for (uint32_t k = lane; k < len; k += 32) dst[k] = __ldg(src + k);
  }
}

// dense[list[i]] = i for the first *d_n entries of the ordered unique list
__global__ void __launch_bounds__(kBlock)
mark_ranked_kernel(uint32_t *__restrict__ dense, const uint32_t *__restrict__ list, uint32_t n_max,
                   const uint32_t *__restrict__ d_n) {
  const uint32_t n = load_count(n_max, d_n);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) dense[__ldg(list + i)] = i;
}

// ranking = head[0, *d_head) ++ tail[0, V - *d_head)
__global__ void __launch_bounds__(kBlock)
concat_rank_kernel(uint32_t *__restrict__ ranking, size_t num_nodes, const uint32_t *__restrict__ head,
                   const uint32_t *__restrict__ d_head, const uint32_t *__restrict__ tail) {
  const size_t h = *d_head < num_nodes ? *d_head : num_nodes;
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < num_nodes; i += stride)
    ranking[i] = i < h ? __ldg(head + i) : __ldg(tail + (i - h));
}

struct HeurWs {
  uint32_t *lens, *offs, *nbr, *pos, *n2o, *out_degree, *deg_rank, *dense, *miss_src, *miss_dst, *cache_src,
      *cache_dst, *num_items, *counts;
  void *table, *chain_ws, *sort_ws, *scan_temp;
  size_t capacity, sort_bytes, scan_bytes, max_unique;
};

size_t scan_temp_bytes(size_t n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n,
                                (cudaStream_t)0);
  return bytes;
}

size_t carve(HeurWs *w, void *base, size_t V, size_t T, size_t N) {
  char *p = (char *)base;
  auto take = [&](size_t bytes) { char *r = p; p += align256(bytes ? bytes : 4); return r; };
  w->max_unique = T + N < V ? T + N : V;
  w->capacity = fgnn_k_ht_capacity(w->max_unique ? w->max_unique : 1);
  w->sort_bytes = fgnn_k_presc_rank_workspace_bytes(V);
  w->scan_bytes = scan_temp_bytes(T);
  w->lens = (uint32_t *)take(T * 4);
  w->offs = (uint32_t *)take(T * 4);
  w->nbr = (uint32_t *)take(N * 4);
  w->pos = (uint32_t *)take(N * 4);
  w->n2o = (uint32_t *)take((w->max_unique + 1) * 4);
  w->out_degree = (uint32_t *)take(V * 4);
  w->deg_rank = (uint32_t *)take(V * 4);
  w->dense = (uint32_t *)take(V * 4);
  w->miss_src = (uint32_t *)take(V * 4);
  w->miss_dst = (uint32_t *)take(V * 4);
  w->cache_src = (uint32_t *)take(V * 4);
  w->cache_dst = (uint32_t *)take(V * 4);
  w->num_items = (uint32_t *)take(4);
  w->counts = (uint32_t *)take(8);
  w->table = take(fgnn_k_ht_bytes(w->capacity));
  w->chain_ws = take(FGNN_CHAIN_WS_BYTES);
  w->sort_ws = take(w->sort_bytes);
  w->scan_temp = take(w->scan_bytes);
  return (size_t)(p - (char *)base);
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" int fgnn_k_row_len_sum(const uint32_t *indptr, const uint32_t *nodes, size_t n,
                                  unsigned long long *d_total, fgnn_stream_t stream) {
  if (!d_total) return FGNN_ERR_BAD_ARG;
  cudaError_t e = cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return 0;
  if (!indptr || !nodes) return FGNN_ERR_BAD_ARG;
  row_len_sum_kernel<<<persistent_grid(n, 4 * kBlock, 8, false, true), kBlock, 0, (cudaStream_t)stream>>>(
      indptr, nodes, n, d_total);
  note_launch();
  return check_last();
}

extern "C" size_t fgnn_k_rank_heuristic_workspace_bytes(size_t num_nodes, size_t num_train,
                                                        size_t num_neighbours) {
  HeurWs w;
  return carve(&w, nullptr, num_nodes, num_train, num_neighbours);
}

extern "C" int fgnn_k_rank_by_heuristic(const uint32_t *indptr, const uint32_t *indices, size_t num_nodes,
                                        size_t num_edges, const uint32_t *train_set, size_t num_train,
                                        size_t num_neighbours, uint32_t *ranking_nodes, void *workspace,
                                        size_t workspace_bytes, fgnn_stream_t stream) {
  if (num_nodes == 0) return 0;
  if (!indptr || !ranking_nodes || !workspace || (num_edges && !indices) || (num_train && !train_set))
    return FGNN_ERR_BAD_ARG;
  if (num_nodes > 0xFFFFFFF0ull || num_train > num_nodes || num_neighbours > 0x7FFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  if (workspace_bytes < fgnn_k_rank_heuristic_workspace_bytes(num_nodes, num_train, num_neighbours))
    return FGNN_ERR_BAD_ARG;
  HeurWs w;
  carve(&w, workspace, num_nodes, num_train, num_neighbours);
  cudaStream_t st = (cudaStream_t)stream;
  const uint32_t V = (uint32_t)num_nodes, T = (uint32_t)num_train, N = (uint32_t)num_neighbours;
  cudaError_t e = cudaMemsetAsync(w.chain_ws, 0, FGNN_CHAIN_WS_BYTES, st);
  if (e != cudaSuccess) return (int)e;
  // 1 + 2: ordered unique list of train_set ++ expanded neighbour rows
  if (int rc = fgnn_k_ht_reset(w.table, w.capacity, w.num_items, stream)) return rc;
  if (T) {
    if (int rc = fgnn_k_ht_fill_unique(w.table, w.capacity, train_set, T, nullptr, w.n2o, w.num_items, stream))
      return rc;
    const int grid_t = persistent_grid(T, 4 * kBlock, 8, false, true);
    row_len_kernel<<<grid_t, kBlock, 0, st>>>(indptr, train_set, T, w.lens);
    e = cub::DeviceScan::ExclusiveSum(w.scan_temp, w.scan_bytes, w.lens, w.offs, (int64_t)T, st);
    if (e != cudaSuccess) return (int)e;
    note_launch(2);
    if (N) {
      expand_rows_kernel<<<persistent_grid(T, kBlock / 32, 8, false, true), kBlock, 0, st>>>(
          indptr, indices, train_set, T, w.offs, w.nbr);
      note_launch();
      if (int rc = fgnn_k_ht_fill_duplicates(w.table, w.capacity, w.nbr, N, nullptr, w.pos, w.n2o, w.num_items,
                                             w.chain_ws, stream))
        return rc;
    }
  }
  // 3: the cache_by_degree ranking, minus what is already ranked, in its own order
  if (int rc = fgnn_k_rank_by_degree(indices, num_edges, num_nodes, w.out_degree, w.deg_rank, w.sort_ws,
                                     w.sort_bytes, stream))
    return rc;
  e = cudaMemsetAsync(w.dense, 0xFF, num_nodes * 4, st);
  if (e != cudaSuccess) return (int)e;
  mark_ranked_kernel<<<persistent_grid(w.max_unique ? w.max_unique : 1, 4 * kBlock, 8, false, true), kBlock, 0, st>>>(
      w.dense, w.n2o, (uint32_t)w.max_unique, w.num_items);
  note_launch();
  if (int rc = fgnn_k_cache_split(w.dense, w.deg_rank, V, nullptr, w.miss_src, w.miss_dst, w.cache_src, w.cache_dst,
                                  w.counts, w.chain_ws, stream))
    return rc;
  concat_rank_kernel<<<persistent_grid(num_nodes, 4 * kBlock, 8, false, true), kBlock, 0, st>>>(
      ranking_nodes, num_nodes, w.n2o, w.num_items, w.miss_src);
  note_launch();
  return check_last();
}
