// Block hand-off in CSC form.  The reference hands every sampled layer to DGL as
// COO (`create_unitgraph_from_coo(2, num_src, num_dst, row, col)`, samgraph/torch/adapter.py:92-95)
// and DGL converts it to CSC on first use inside the trainer — the "convert" stage the reference
// times as kLogL1ConvertTime.  Its DGL patch already exposes `create_unitgraph_from_csc`
// (3rdparty/dgl.patch:30-57: indptr over dst nodes, indices = src ids, edge_ids); this file builds
// exactly those three arrays on the GPU from the TrainGraph's (row, col).
//
//   indptr[d]   = number of edges with col < d            (d in [0, num_dst])
//   indices[i]  = row[perm[i]]
//   edge_ids[i] = perm[i],  perm = stable sort of the edge ids by col
//
// khop0 / khop2 / hash-dedup / random-walk blocks are emitted seed-major (col ascending), so perm is
// the identity and only the indptr is computed (one binary search per dst node).  khop1 / weighted
// blocks are ordered by the seeds' GLOBAL ids (the reference's sort-by-src), so their local col is
// a permutation of runs: those take one radix sort of (col, edge id) pairs over log2(num_dst) bits.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace fgnn {
namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

inline int key_bits(uint32_t num_dst_max) {  // keys are in [0, num_dst_max]; num_dst_max = padding
  int b = 1;
  while (b < 32 && (num_dst_max >> b) != 0) ++b;
  return b;
}

size_t sort_temp_bytes(uint32_t e_max, int bits) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)e_max, 0, bits,
                                  (cudaStream_t)0);
  return bytes;
}

// first index in col[0, n) whose value is >= d (col ascending)
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t *__restrict__ col, uint32_t n, uint32_t d) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(col + mid) < d) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kBlock)
csc_indptr_kernel(const uint32_t *__restrict__ col, uint32_t e_max, const uint32_t *__restrict__ d_e,
                  uint32_t num_dst, uint32_t *__restrict__ indptr) {
  const uint32_t e = load_count(e_max, d_e);
  for (uint32_t d = blockIdx.x * kBlock + threadIdx.x; d <= num_dst; d += gridDim.x * kBlock)
    indptr[d] = lower_bound_u32(col, e, d);
}

// sorted path, only when the caller wants its own copies: indices = row, edge_ids = 0..e-1 (live entries only)
__global__ void __launch_bounds__(kBlock)
csc_identity_kernel(const uint32_t *__restrict__ row, uint32_t e_max, const uint32_t *__restrict__ d_e,
                    uint32_t *__restrict__ indices, uint32_t *__restrict__ edge_ids) {
  const uint32_t e = load_count(e_max, d_e);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < e; i += gridDim.x * kBlock) {
    if (indices) indices[i] = __ldg(row + i);
    if (edge_ids) edge_ids[i] = i;
  }
}

// unsorted path: keys = col (entries past the live count get the padding key num_dst_max, which sorts last)
__global__ void __launch_bounds__(kBlock)
csc_keys_kernel(const uint32_t *__restrict__ col, uint32_t e_max, const uint32_t *__restrict__ d_e,
                uint32_t pad_key, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  const uint32_t e = load_count(e_max, d_e);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < e_max; i += gridDim.x * kBlock) {
    keys[i] = i < e ? __ldg(col + i) : pad_key;
    vals[i] = i;
  }
}

__global__ void __launch_bounds__(kBlock)
csc_permute_kernel(const uint32_t *__restrict__ row, const uint32_t *__restrict__ perm, uint32_t e_max,
                   const uint32_t *__restrict__ d_e, uint32_t *__restrict__ indices,
                   uint32_t *__restrict__ edge_ids) {
  const uint32_t e = load_count(e_max, d_e);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < e; i += gridDim.x * kBlock) {
    const uint32_t p = __ldg(perm + i);
    indices[i] = __ldg(row + p);
    if (edge_ids) edge_ids[i] = p;
  }
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_coo_to_csc_workspace_bytes(uint32_t e_max, uint32_t num_dst_max) {
  return 4 * align256((size_t)e_max * 4) + align256(sort_temp_bytes(e_max, key_bits(num_dst_max)));
}

extern "C" int fgnn_k_coo_to_csc(const uint32_t *row, const uint32_t *col, uint32_t e_max, const uint32_t *d_e,
                                 uint32_t num_dst, int col_sorted, uint32_t *indptr, uint32_t *indices,
                                 uint32_t *edge_ids, void *workspace, size_t workspace_bytes,
                                 fgnn_stream_t stream) {
  if (!indptr) return FGNN_ERR_BAD_ARG;
  if (e_max && (!row || !col)) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid_d = persistent_grid((uint64_t)num_dst + 1, kBlock, 8, false);
  if (e_max == 0) {
    csc_indptr_kernel<<<grid_d, kBlock, 0, st>>>(col, 0, nullptr, num_dst, indptr);
    note_launch();
    return check_last();
  }
  const int grid_e = persistent_grid(e_max, 2 * kBlock, 8, false);
  if (col_sorted) {
    // identity permutation: `indices` (when asked for) is a copy of row, edge_ids = 0..e-1
    csc_indptr_kernel<<<grid_d, kBlock, 0, st>>>(col, e_max, d_e, num_dst, indptr);
    note_launch();
    if (indices == row) indices = nullptr;
    if (edge_ids || indices) {
      csc_identity_kernel<<<grid_e, kBlock, 0, st>>>(row, e_max, d_e, indices, edge_ids);
      note_launch();
    }
    return check_last();
  }
  if (!indices || !workspace) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_coo_to_csc_workspace_bytes(e_max, num_dst)) return FGNN_ERR_BAD_ARG;
  const size_t nb = align256((size_t)e_max * 4);
  char *p = (char *)workspace;
  uint32_t *keys_in = (uint32_t *)p, *keys_out = (uint32_t *)(p + nb);
  uint32_t *vals_in = (uint32_t *)(p + 2 * nb), *vals_out = (uint32_t *)(p + 3 * nb);
  void *temp = p + 4 * nb;
  size_t temp_bytes = workspace_bytes - 4 * nb;
  const int bits = key_bits(num_dst);
  csc_keys_kernel<<<grid_e, kBlock, 0, st>>>(col, e_max, d_e, num_dst, keys_in, vals_in);
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                  (int64_t)e_max, 0, bits, st);  // LSD radix sort: stable
  if (e != cudaSuccess) return (int)e;
  csc_indptr_kernel<<<grid_d, kBlock, 0, st>>>(keys_out, e_max, d_e, num_dst, indptr);
  csc_permute_kernel<<<grid_e, kBlock, 0, st>>>(row, vals_out, e_max, d_e, indices, edge_ids);
  note_launch(3);
  return check_last();
}
