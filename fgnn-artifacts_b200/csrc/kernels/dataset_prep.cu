// Dataset preparation on the GPU: the tables the weighted samplers and the
// non-PreSC cache policies read, which the reference builds offline on the CPU with
// utility/data-process/toolkit/{weight,cache}/*.cc and loads from *.bin files.
//
//   alias table   create_alias_table.cc:96-180      (Vose with two FIFO queues, per CSR row)
//   prefix table  create_prob_prefix_table.cc:94-123 (running fp32 sum per row)
//   out degree    common/graph_loader.cc:109-147
//   degree rank   toolkit/cache/cache_by_degree.cc:36-58  (sort {out_degree, id} descending)
//
// Both weight tables are order-sensitive fp32 recurrences inside a row (the sums are sequential and the
// FIFO order decides which large entry pays for which small one), so a row is walked by ONE thread exactly
// like the reference's loop body; parallelism is across rows.  Rows are handed out in blocks of 32 by a
// global ticket so that a warp stuck on hub rows does not hold back the rest of the grid.
#include "common.cuh"

namespace fgnn {
namespace {

constexpr uint32_t kRowsPerTicket = 32;  // one row per lane

__device__ unsigned long long g_prep_ticket[2];  // {next block of rows, finished CTAs}

template <typename F>
__device__ __forceinline__ void for_each_row_dynamic(size_t num_nodes, F &&body) {
  const uint32_t lane = threadIdx.x & 31;
  for (;;) {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(&g_prep_ticket[0], 1ull);
    t = __shfl_sync(0xFFFFFFFFu, t, 0);
    const unsigned long long first = t * kRowsPerTicket;
    if (first >= num_nodes) break;
    const unsigned long long v = first + lane;
    if (v < num_nodes) body((size_t)v);
    __syncwarp();
  }
  // the last CTA to leave re-arms the ticket for the next launch
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&g_prep_ticket[1], 1ull) == gridDim.x - 1) {
      g_prep_ticket[0] = 0;
      g_prep_ticket[1] = 0;
      __threadfence();
    }
  }
}

// create_alias_table.cc:101-176.  prob_table doubles as the row's scaled-weight array w[]: the reference's
// `prob_table[off+small] = weights[small]` then stores the value already in place, and every index still
// queued at the end gets 1.  queue = u32[2*E]: row v owns [2*off, 2*off+2*len): smalls ring then larges ring,
// each of capacity len (an index is in at most one queue at a time).
__global__ void __launch_bounds__(kBlock)
alias_table_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                   size_t num_nodes, const float *__restrict__ weights, float *prob,
                   uint32_t *alias, uint32_t *queue) {
  for_each_row_dynamic(num_nodes, [&](size_t v) {
    const uint32_t off = indptr[v];
    const uint32_t len = indptr[v + 1] - off;
    if (len == 0) return;
    float *w = prob + off;
    const uint32_t *nb = indices + off;
    uint32_t *al = alias + off;
    uint32_t *smalls = queue + 2 * (size_t)off, *larges = smalls + len;
    float sum = 0.0f;
    for (uint32_t i = 0; i < len; ++i) sum = __fadd_rn(sum, weights[off + i]);        // :112-126
    uint32_t sh = 0, sn = 0, lh = 0, ln = 0;  // ring heads and live counts
    const float flen = (float)len;
    for (uint32_t i = 0; i < len; ++i) {                                                // :128-141
      const float x = __fmul_rn(__fdiv_rn(weights[off + i], sum), flen);
      w[i] = x;
      al[i] = 0;  // entries whose prob ends at 1 keep the reference's zero-initialised alias (create_alias_table.cc:210)
      if (x < 1.0f) { smalls[sn++] = i; } else { larges[ln++] = i; }
    }
    uint32_t st = sn == len ? 0 : sn, lt = ln == len ? 0 : ln;  // ring tails
    while (sn != 0 && ln != 0) {                                                        // :145-161
      const uint32_t s = smalls[sh], l = larges[lh];
      sh = sh + 1 == len ? 0 : sh + 1; --sn;
      lh = lh + 1 == len ? 0 : lh + 1; --ln;
      al[s] = nb[l];
      const float wl = __fsub_rn(w[l], __fsub_rn(1.0f, w[s]));
      w[l] = wl;
      if (wl < 1.0f) { smalls[st] = l; st = st + 1 == len ? 0 : st + 1; ++sn; }
      else { larges[lt] = l; lt = lt + 1 == len ? 0 : lt + 1; ++ln; }
    }
    while (ln != 0) { w[larges[lh]] = 1.0f; lh = lh + 1 == len ? 0 : lh + 1; --ln; }    // :163-168
    while (sn != 0) { w[smalls[sh]] = 1.0f; sh = sh + 1 == len ? 0 : sh + 1; --sn; }    // :170-175
  });
}

// create_prob_prefix_table.cc:99-120
__global__ void __launch_bounds__(kBlock)
prefix_table_kernel(const uint32_t *__restrict__ indptr, size_t num_nodes,
                    const float *__restrict__ weights, float *prefix) {
  for_each_row_dynamic(num_nodes, [&](size_t v) {
    const uint32_t off = indptr[v];
    const uint32_t len = indptr[v + 1] - off;
    float sum = 0.0f;
    for (uint32_t i = 0; i < len; ++i) {
      sum = __fadd_rn(sum, weights[off + i]);
      prefix[off + i] = sum;
    }
  });
}

// graph_loader.cc:126-137: out_degree[indices[e]]++ over all edges (E can exceed 2^31: 64-bit indexing)
__global__ void __launch_bounds__(kBlock)
out_degree_kernel(const uint32_t *__restrict__ indices, size_t num_edges, uint32_t *out_degree) {
  const size_t stride = (size_t)gridDim.x * kBlock * 4;
  for (size_t base = ((size_t)blockIdx.x * kBlock + threadIdx.x) * 4; base < num_edges; base += stride) {
    if (base + 4 <= num_edges && (reinterpret_cast<uintptr_t>(indices + base) & 15) == 0) {
      const uint4 q = *reinterpret_cast<const uint4 *>(indices + base);
      atomicAdd(out_degree + q.x, 1u);
      atomicAdd(out_degree + q.y, 1u);
      atomicAdd(out_degree + q.z, 1u);
      atomicAdd(out_degree + q.w, 1u);
    } else {
      for (size_t e = base; e < num_edges && e < base + 4; ++e) atomicAdd(out_degree + indices[e], 1u);
    }
  }
}

__global__ void __launch_bounds__(kBlock) iota_kernel(uint32_t *out, size_t n) {
  const size_t stride = (size_t)gridDim.x * kBlock;
  for (size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_alias_table_workspace_bytes(size_t num_edges) {
  return 2 * num_edges * sizeof(uint32_t);
}

extern "C" int fgnn_k_build_alias_table(const uint32_t *indptr, const uint32_t *indices, size_t num_nodes,
                                        size_t num_edges, const float *weights, float *prob_table,
                                        uint32_t *alias_table, void *workspace, size_t workspace_bytes,
                                        fgnn_stream_t stream) {
  if (num_nodes == 0 || num_edges == 0) return 0;
  if (!indptr || !indices || !weights || !prob_table || !alias_table || !workspace) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_alias_table_workspace_bytes(num_edges)) return FGNN_ERR_BAD_ARG;
  static const int occ = occupancy(alias_table_kernel, kBlock, 0);
  const int grid = persistent_grid((num_nodes + kRowsPerTicket - 1) / kRowsPerTicket, kBlock / 32, occ, false, true);
  alias_table_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(indptr, indices, num_nodes, weights, prob_table,
                                                                alias_table, (uint32_t *)workspace);
  note_launch();
  return check_last();
}

extern "C" int fgnn_k_build_prefix_table(const uint32_t *indptr, size_t num_nodes, const float *weights,
                                         float *prob_prefix_table, fgnn_stream_t stream) {
  if (num_nodes == 0) return 0;
  if (!indptr || !weights || !prob_prefix_table) return FGNN_ERR_BAD_ARG;
  static const int occ = occupancy(prefix_table_kernel, kBlock, 0);
  const int grid = persistent_grid((num_nodes + kRowsPerTicket - 1) / kRowsPerTicket, kBlock / 32, occ, false, true);
  prefix_table_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(indptr, num_nodes, weights, prob_prefix_table);
  note_launch();
  return check_last();
}

extern "C" int fgnn_k_out_degree(const uint32_t *indices, size_t num_edges, uint32_t *out_degree,
                                 size_t num_nodes, fgnn_stream_t stream) {
  if (!out_degree) return FGNN_ERR_BAD_ARG;
  cudaError_t e = cudaMemsetAsync(out_degree, 0, num_nodes * sizeof(uint32_t), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  if (num_edges == 0) return 0;
  if (!indices) return FGNN_ERR_BAD_ARG;
  static const int occ = occupancy(out_degree_kernel, kBlock, 0);
  const int grid = persistent_grid((num_edges + 3) / 4, kBlock * 4, occ, false, true);
  out_degree_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(indices, num_edges, out_degree);
  note_launch();
  return check_last();
}

// cache_by_degree.cc:36-58 == the PreSC ranking with freq := out_degree (same {key, id} descending order)
extern "C" int fgnn_k_rank_by_degree(const uint32_t *indices, size_t num_edges, size_t num_nodes,
                                     uint32_t *out_degree, uint32_t *ranking_nodes, void *workspace,
                                     size_t workspace_bytes, fgnn_stream_t stream) {
  if (int rc = fgnn_k_out_degree(indices, num_edges, out_degree, num_nodes, stream)) return rc;
  return fgnn_k_presc_rank(out_degree, num_nodes, ranking_nodes, workspace, workspace_bytes, stream);
}

extern "C" size_t fgnn_k_rank_random_workspace_bytes(size_t num_nodes) {
  return align256(num_nodes * sizeof(uint32_t)) + fgnn_k_shuffle_workspace_bytes(num_nodes);
}

extern "C" int fgnn_k_rank_random(size_t num_nodes, uint64_t seed, uint32_t *ranking_nodes, void *workspace,
                                  size_t workspace_bytes, fgnn_stream_t stream) {
  if (num_nodes == 0) return 0;
  if (!ranking_nodes || !workspace || num_nodes > 0xFFFFFFFFull) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_rank_random_workspace_bytes(num_nodes)) return FGNN_ERR_BAD_ARG;
  uint32_t *ids = (uint32_t *)workspace;
  const size_t ib = align256(num_nodes * sizeof(uint32_t));
  iota_kernel<<<persistent_grid(num_nodes, 4 * kBlock, 8, false, true), kBlock, 0, (cudaStream_t)stream>>>(ids, num_nodes);
  note_launch();
  // epoch word 0xCAC4E: keeps the ranking's Philox stream apart from the per-epoch train-set shuffles
  return fgnn_k_shuffle(ids, num_nodes, seed, 0xCAC4Eull << 32, ranking_nodes, (char *)workspace + ib,
                        workspace_bytes - ib, stream);
}
