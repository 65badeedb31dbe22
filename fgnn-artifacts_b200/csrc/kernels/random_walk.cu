// PinSAGE neighbour sampling: W random walks of length L per start node, then
// the K most-visited nodes per start node.
//
// Reference: sample_random_walk (cuda_sampling_random_walk.cu:43-109) writes all
// S*W*L (start, visited) pairs to HBM; FrequencyHashmap::GetTopK
// (cuda_frequency_hashmap.cu:1143-1367) then runs 11 synchronised steps over a
// 64-bucket x 12-byte *global-memory* hash table per start node plus a 64-bit
// radix sort of all unique pairs.
//
// Here one launch does both: a tile's walks are advanced by one thread per walk
// (the dependent indptr -> indices loads of many walks overlap), the W*L visited
// ids of a node are handed to a 2^k-lane group through shared memory, and the
// frequency count / ranking is done with warp shuffles: count_p = #lanes with
// the same id, representative = first occurrence, rank = #representatives that
// beat it on (count desc, first-occurrence asc) — exactly the order produced by
// the reference's stable descending sort on (node-major, count).  Only the final
// (start, visited, count) triples reach HBM.
#include "common.cuh"

namespace fgnn {
namespace {

struct RwSmem {
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

struct RwWs {
  uint32_t *pad_dst, *pad_cnt, *node_cnt;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(RwWs *w, void *base, uint32_t n_max, uint32_t K) {
  char *p = (char *)base;
  w->pad_dst = (uint32_t *)p; p += align256((size_t)n_max * K * 4);
  w->pad_cnt = (uint32_t *)p; p += align256((size_t)n_max * K * 4);
  w->node_cnt = (uint32_t *)p; p += align256((size_t)n_max * 4);
  return (size_t)(p - (char *)base);
}

template <int G /* lanes per node, power of two >= W*L */>
__global__ void __launch_bounds__(kBlock)
random_walk_topk_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                        const uint32_t *__restrict__ input, uint32_t n_max,
                        const uint32_t *__restrict__ d_n, uint32_t L, double restart_prob, uint32_t W,
                        uint32_t K, RngKey key, uint32_t *__restrict__ out_src,
                        uint32_t *__restrict__ out_dst, uint32_t *__restrict__ out_src_local,
                        uint32_t *__restrict__ out_data, uint32_t *d_num_out,
                        uint32_t *__restrict__ tmp_src, uint32_t *__restrict__ tmp_dst, RwWs wsb,
                        uint32_t NT /* nodes per tile */, ChainWs *ws) {
  extern __shared__ uint32_t s_vis[];  // [NT][G] visited id or EMPTY
  __shared__ RwSmem sm;
  const uint32_t EPN = W * L;
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, NT, &begin, &end);

  unsigned long long partial = 0;
  for (uint32_t t0 = begin; t0 < end; t0 += NT) {
    // ---- walks: one thread per (node, walk) ---------------------------------
    for (uint32_t e = threadIdx.x; e < NT * G; e += kBlock) s_vis[e] = kEmpty;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < NT * W; t += kBlock) {
      const uint32_t nl = t / W, w = t - nl * W;
      const uint32_t node_idx = t0 + nl;
      if (node_idx >= end) continue;
      const uint32_t start = __ldg(input + node_idx);
      const uint32_t item = node_idx * W + w;
      uint32_t node = start;
      for (uint32_t s = 0; s < L; ++s) {
        const uint32_t lpos = s * W + w;  // random_walk.cu:77-78
        uint32_t visited = kEmpty;
        if (node != kEmpty) {
          const uint32_t off = __ldg(indptr + node);
          const uint32_t len = __ldg(indptr + node + 1) - off;
          if (len == 0) {
            node = kEmpty;
          } else {
            // draws 3s, 3s+1, 3s+2
            const uint32_t r0 = rand_u32(key, item, 3 * s);
            const uint32_t r1 = rand_u32(key, item, 3 * s + 1);
            const uint32_t r2 = rand_u32(key, item, 3 * s + 2);
            visited = __ldg(indices + (size_t)off + (r0 % len));
            node = visited;
            if (uniform_f64(r1, r2) < restart_prob) node = kEmpty;
          }
        }
        s_vis[nl * G + lpos] = visited;
        if (tmp_src) {
          const size_t gpos = (size_t)node_idx * EPN + lpos;
          tmp_src[gpos] = visited == kEmpty ? kEmpty : start;
          tmp_dst[gpos] = visited;
        }
      }
    }
    __syncthreads();

    // ---- frequency top-K: G lanes per node ------------------------------------
    for (uint32_t g0 = 0; g0 < NT; g0 += kBlock / G) {  // uniform trip count: shuffles below
      const uint32_t g = g0 + threadIdx.x / G;
      const uint32_t node_idx = (g < NT) ? t0 + g : end;
      const uint32_t pl = threadIdx.x & (G - 1);
      const uint32_t d = (g < NT) ? s_vis[g * G + pl] : kEmpty;
      const bool valid = (d != kEmpty) && (node_idx < end);
      uint32_t cnt = 0;
      bool first = true;
#pragma unroll
      for (int q = 0; q < G; ++q) {
        const uint32_t dq = __shfl_sync(0xFFFFFFFFu, d, q, G);
        if (dq == d) {
          ++cnt;
          if ((uint32_t)q < pl) first = false;
        }
      }
      const bool rep = valid && first;
      uint32_t rank = 0;
#pragma unroll
      for (int q = 0; q < G; ++q) {
        const uint32_t cq = __shfl_sync(0xFFFFFFFFu, cnt, q, G);
        const int rq = __shfl_sync(0xFFFFFFFFu, (int)rep, q, G);
        if (rq && (cq > cnt || (cq == cnt && (uint32_t)q < pl))) ++rank;
      }
      const uint32_t repmask = __ballot_sync(0xFFFFFFFFu, rep);
      const uint32_t gshift = (threadIdx.x & 31) & ~(G - 1);
      const uint32_t gmask = (G == 32) ? 0xFFFFFFFFu : (((1u << G) - 1u) << gshift);
      uint32_t nrep = __popc(repmask & gmask);
      if (nrep > K) nrep = K;
      if (node_idx < end) {
        if (rep && rank < K) {
          wsb.pad_dst[(size_t)node_idx * K + rank] = d;
          wsb.pad_cnt[(size_t)node_idx * K + rank] = cnt;
        }
        if (pl == 0) {
          wsb.node_cnt[node_idx] = nrep;
          partial += nrep;
        }
      }
    }
    __syncthreads();
  }

  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);

  // ---- compact (node-major): compact_output_revised, frequency_hashmap.cu:644-676
  for (uint32_t t0 = begin; t0 < end; t0 += kBlock) {
    const uint32_t node_idx = t0 + threadIdx.x;
    const uint32_t c = node_idx < end ? wsb.node_cnt[node_idx] : 0u;
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(c, sm.warp, &tile_total);
    if (c) {
      const uint32_t start = __ldg(input + node_idx);
      for (uint32_t r = 0; r < c; ++r) {
        const size_t o = (size_t)base + excl + r;
        out_dst[o] = wsb.pad_dst[(size_t)node_idx * K + r];
        out_data[o] = wsb.pad_cnt[(size_t)node_idx * K + r];
        if (out_src) out_src[o] = start;
        if (out_src_local) out_src_local[o] = node_idx;
      }
    }
    base += tile_total;
  }
  chain_finish(ws, &sm.chain);
}

// The same launch for W*L > 32 visited ids per start node ("top-k over a frequency hashmap in shared memory"):
// the ids of a node stay in shared memory, one WARP per node counts them (count_p = #positions with the same id,
// representative = first occurrence) and ranks the representatives by (count desc, first occurrence asc) with two
// O(EPN^2 / 32) passes over shared memory — the order of the reference's stable descending sort, as above.
__global__ void __launch_bounds__(kBlock)
random_walk_topk_big_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                            const uint32_t *__restrict__ input, uint32_t n_max,
                            const uint32_t *__restrict__ d_n, uint32_t L, double restart_prob, uint32_t W,
                            uint32_t K, RngKey key, uint32_t *__restrict__ out_src,
                            uint32_t *__restrict__ out_dst, uint32_t *__restrict__ out_src_local,
                            uint32_t *__restrict__ out_data, uint32_t *d_num_out,
                            uint32_t *__restrict__ tmp_src, uint32_t *__restrict__ tmp_dst, RwWs wsb,
                            uint32_t NT /* nodes per tile */, ChainWs *ws) {
  extern __shared__ uint32_t s_dyn[];
  __shared__ RwSmem sm;
  const uint32_t EPN = W * L;
  uint32_t *s_vis = s_dyn;             // [NT][EPN] visited id or EMPTY
  uint32_t *s_cnt = s_dyn + NT * EPN;  // [NT][EPN] count at the representative position, else 0
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, NT, &begin, &end);
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  unsigned long long partial = 0;
  for (uint32_t t0 = begin; t0 < end; t0 += NT) {
    for (uint32_t e = threadIdx.x; e < NT * EPN; e += kBlock) s_vis[e] = kEmpty;
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < NT * W; t += kBlock) {  // walks: identical to the kernel above
      const uint32_t nl = t / W, w = t - nl * W;
      const uint32_t node_idx = t0 + nl;
      if (node_idx >= end) continue;
      const uint32_t start = __ldg(input + node_idx);
      const uint32_t item = node_idx * W + w;
      uint32_t node = start;
      for (uint32_t s = 0; s < L; ++s) {
        const uint32_t lpos = s * W + w;
        uint32_t visited = kEmpty;
        if (node != kEmpty) {
          const uint32_t off = __ldg(indptr + node);
          const uint32_t len = __ldg(indptr + node + 1) - off;
          if (len == 0) {
            node = kEmpty;
          } else {
            const uint32_t r0 = rand_u32(key, item, 3 * s);
            const uint32_t r1 = rand_u32(key, item, 3 * s + 1);
            const uint32_t r2 = rand_u32(key, item, 3 * s + 2);
            visited = __ldg(indices + (size_t)off + (r0 % len));
            node = visited;
            if (uniform_f64(r1, r2) < restart_prob) node = kEmpty;
          }
        }
        s_vis[nl * EPN + lpos] = visited;
        if (tmp_src) {
          const size_t gpos = (size_t)node_idx * EPN + lpos;
          tmp_src[gpos] = visited == kEmpty ? kEmpty : start;
          tmp_dst[gpos] = visited;
        }
      }
    }
    __syncthreads();
    for (uint32_t g = warp; g < NT; g += kBlock / 32) {  // one warp per node
      const uint32_t node_idx = t0 + g;
      if (node_idx >= end) continue;
      const uint32_t *vis = s_vis + g * EPN;
      uint32_t *cntv = s_cnt + g * EPN;
      for (uint32_t pl = lane; pl < EPN; pl += 32) {
        const uint32_t d = vis[pl];
        uint32_t cnt = 0;
        bool first = true;
        for (uint32_t q = 0; q < EPN; ++q) {
          if (vis[q] == d) {
            ++cnt;
            if (q < pl) first = false;
          }
        }
        cntv[pl] = (d != kEmpty && first) ? cnt : 0u;
      }
      __syncwarp();
      uint32_t nrep = 0;
      for (uint32_t pl = lane; pl < EPN; pl += 32) {
        const uint32_t c = cntv[pl];
        if (!c) continue;
        ++nrep;
        uint32_t rank = 0;
        for (uint32_t q = 0; q < EPN; ++q) {
          const uint32_t cq = cntv[q];
          if (cq > c || (cq == c && q < pl)) ++rank;
        }
        if (rank < K) {
          wsb.pad_dst[(size_t)node_idx * K + rank] = vis[pl];
          wsb.pad_cnt[(size_t)node_idx * K + rank] = c;
        }
      }
#pragma unroll
      for (int dlt = 16; dlt > 0; dlt >>= 1) nrep += __shfl_down_sync(0xFFFFFFFFu, nrep, dlt);
      if (lane == 0) {
        if (nrep > K) nrep = K;
        wsb.node_cnt[node_idx] = nrep;
        partial += nrep;
      }
    }
    __syncthreads();
  }

  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);
  for (uint32_t t0 = begin; t0 < end; t0 += kBlock) {
    const uint32_t node_idx = t0 + threadIdx.x;
    const uint32_t c = node_idx < end ? wsb.node_cnt[node_idx] : 0u;
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(c, sm.warp, &tile_total);
    if (c) {
      const uint32_t start = __ldg(input + node_idx);
      for (uint32_t r = 0; r < c; ++r) {
        const size_t o = (size_t)base + excl + r;
        out_dst[o] = wsb.pad_dst[(size_t)node_idx * K + r];
        out_data[o] = wsb.pad_cnt[(size_t)node_idx * K + r];
        if (out_src) out_src[o] = start;
        if (out_src_local) out_src_local[o] = node_idx;
      }
    }
    base += tile_total;
  }
  chain_finish(ws, &sm.chain);
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_sample_random_walk_workspace_bytes(uint32_t n_max, uint32_t K) {
  RwWs w;
  return carve(&w, nullptr, n_max ? n_max : 1, K ? K : 1);
}

extern "C" int fgnn_k_sample_random_walk(const uint32_t *indptr, const uint32_t *indices,
                                         const uint32_t *input, uint32_t n_max, const uint32_t *d_n,
                                         uint32_t walk_len, double restart_prob, uint32_t num_walk,
                                         uint32_t K, fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst,
                                         uint32_t *out_src_local, uint32_t *out_data,
                                         uint32_t *d_num_out, uint32_t *tmp_src, uint32_t *tmp_dst,
                                         void *workspace, size_t workspace_bytes, void *chain_ws,
                                         fgnn_stream_t stream) {
  if (!indptr || !indices || !out_dst || !out_data || !d_num_out || !chain_ws) return FGNN_ERR_BAD_ARG;
  if ((tmp_src == nullptr) != (tmp_dst == nullptr)) return FGNN_ERR_BAD_ARG;
  const uint32_t EPN = walk_len * num_walk;
  if (walk_len == 0 || num_walk == 0 || K == 0) return FGNN_ERR_BAD_ARG;
  if (EPN > 4096) return FGNN_ERR_UNSUPPORTED;  // shared-memory path: 8 nodes x EPN x 8 bytes per tile
  if ((uint64_t)n_max * (K > EPN ? K : EPN) > 0xFFFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_max == 0) return (int)cudaMemsetAsync(d_num_out, 0, sizeof(uint32_t), st);
  if (!input || !workspace) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_sample_random_walk_workspace_bytes(n_max, K)) return FGNN_ERR_BAD_ARG;
  RwWs w;
  carve(&w, workspace, n_max, K);
  uint32_t G = 1;
  while (G < EPN) G <<= 1;
  uint32_t NT = kBlock / num_walk;  // one thread per walk in a tile
  if (NT > (uint32_t)kBlock) NT = kBlock;
  if (NT < 8) NT = 8;
  const size_t smem = (size_t)NT * G * sizeof(uint32_t);
  const RngKey key = make_rng_key(rng);
  if (EPN > 32) {
    if (NT > 64) NT = 64;
    while (NT > 8 && (size_t)NT * EPN * 8 > 96 * 1024) NT >>= 1;
    const size_t smem_big = (size_t)NT * EPN * 8;
    if (smem_big > 200 * 1024) return FGNN_ERR_UNSUPPORTED;
    if (smem_big > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(random_walk_topk_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem_big);
      if (e != cudaSuccess) return (int)e;
    }
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, random_walk_topk_big_kernel, kBlock, smem_big);
    if (occ < 1) occ = 1;
    const int grid = persistent_grid(n_max, NT, occ, true);
    random_walk_topk_big_kernel<<<grid, kBlock, smem_big, st>>>(
        indptr, indices, input, n_max, d_n, walk_len, restart_prob, num_walk, K, key, out_src, out_dst,
        out_src_local, out_data, d_num_out, tmp_src, tmp_dst, w, NT, (ChainWs *)chain_ws);
    note_launch();
    return check_last();
  }
#define FGNN_RW(GG)                                                                                \
  {                                                                                                \
    int occ = 1;                                                                                   \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, random_walk_topk_kernel<GG>, kBlock, smem); \
    if (occ < 1) occ = 1;                                                                          \
    const int grid = persistent_grid(n_max, NT, occ, true);                                        \
    random_walk_topk_kernel<GG><<<grid, kBlock, smem, st>>>(                                       \
        indptr, indices, input, n_max, d_n, walk_len, restart_prob, num_walk, K, key, out_src,     \
        out_dst, out_src_local, out_data, d_num_out, tmp_src, tmp_dst, w, NT, (ChainWs *)chain_ws); \
  }
  switch (G) {
    case 1: FGNN_RW(1); break;
    case 2: FGNN_RW(2); break;
    case 4: FGNN_RW(4); break;
    case 8: FGNN_RW(8); break;
    case 16: FGNN_RW(16); break;
    default: FGNN_RW(32); break;
  }
#undef FGNN_RW
  note_launch();
  return check_last();
}
