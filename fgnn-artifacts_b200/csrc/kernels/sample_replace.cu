// With-replacement samplers: khop1 (uniform), weighted alias, weighted prefix;
// and the rejection-based weighted sampler without replacement (hash dedup).
//
// Reference flow for khop1 / weighted (cuda_sampling_khop1.cu:130-234,
// cuda_sampling_weighted_khop.cu:132-236, ..._prefix.cu:137-255):
//   sample S*f (src,dst) pairs -> radix-sort ALL pairs by src -> flag entries that
//   differ from their successor -> DeviceScan over S*f+1 -> compact.
// Seeds of one layer are unique, so sorting the pairs by src is a permutation of
// the *rows*.  Here only the S seed ids are sorted (f-times less sort work), the
// draws are evaluated edge-parallel straight into rank-major order, and the
// flag+scan+compact is one chunk-chained launch.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace fgnn {
namespace {

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct ReplaceWs {
  uint32_t *keys_in, *keys_out, *vals_in, *vals_out, *tmp_dst;
  void *cub_temp;
  size_t cub_bytes;
};

size_t sort_temp_bytes(uint32_t n_max) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                  (const uint32_t *)nullptr, (uint32_t *)nullptr, (int64_t)n_max, 0, 32,
                                  (cudaStream_t)0);
  return bytes;
}

size_t carve(ReplaceWs *w, void *base, uint32_t n_max, uint32_t fanout) {
  char *p = (char *)base;
  const size_t nb = align256((size_t)n_max * 4);
  w->keys_in = (uint32_t *)p; p += nb;
  w->keys_out = (uint32_t *)p; p += nb;
  w->vals_in = (uint32_t *)p; p += nb;
  w->vals_out = (uint32_t *)p; p += nb;
  w->tmp_dst = (uint32_t *)p; p += align256((size_t)n_max * fanout * 4 + 4);
  w->cub_bytes = sort_temp_bytes(n_max);
  w->cub_temp = p; p += align256(w->cub_bytes);
  return (size_t)(p - (char *)base);
}

// ---------------------------------------------------------------------------
// Ordering the seeds by id WITHOUT a sort (round 2; replaces the cub::DeviceRadixSort of round 1 on the
// khop1 / weighted path, BASELINE config #5).  The seeds of a layer are unique, so the position of seed v in
// ascending id order is simply the number of seeds with a smaller id: mark the seeds in a V-bit bitmap, keep a
// count per 1024-bit block, scan the block counts (chained scan, common.cuh), and
//     rank(v) = blockpre[v >> 10] + popcount(bitmap words of the block below v's word) + popcount(low bits).
// The bitmap and the block counts live in `rank_ws`, which is all-zero between calls (the last kernel clears
// exactly the words it set), so nothing of size V is touched per mini-batch.
// ---------------------------------------------------------------------------
struct RankWs {
  uint32_t *bitmap;    // ceil(V / 32) words
  uint32_t *blocksum;  // nb + 1 words, nb = ceil(V / 1024)
  uint32_t *blockpre;  // nb + 1 words (scratch, no invariant)
  uint32_t *counters;  // [0] = cursor of the rows that produce nothing
  uint32_t nb;
};

size_t carve_rank(RankWs *w, void *base, size_t num_nodes) {
  char *p = (char *)base;
  const size_t words = (num_nodes + 31) / 32, nb = (num_nodes + 1023) / 1024;
  w->bitmap = (uint32_t *)p; p += align256(words * 4);
  w->blocksum = (uint32_t *)p; p += align256((nb + 1) * 4);
  w->counters = (uint32_t *)p; p += 256;
  w->blockpre = (uint32_t *)p; p += align256((nb + 1) * 4);
  w->nb = (uint32_t)nb;
  return (size_t)(p - (char *)base);
}

// sort key = seed id, EMPTY for rows that produce nothing (len == 0, padding); with a rank workspace the live
// seeds are marked in the bitmap at the same time
__global__ void __launch_bounds__(kBlock)
replace_keys_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ input,
                    uint32_t n_max, const uint32_t *__restrict__ d_n, uint32_t *keys, uint32_t *vals,
                    uint32_t *bitmap, uint32_t *blocksum) {
  const uint32_t n = load_count(n_max, d_n);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n_max; i += gridDim.x * kBlock) {
    uint32_t key = kEmpty;
    if (i < n) {
      const uint32_t v = __ldg(input + i);
      if (__ldg(indptr + v + 1) != __ldg(indptr + v)) key = v;
    }
    keys[i] = key;
    vals[i] = i;
    if (bitmap && key != kEmpty) {
      atomicOr(bitmap + (key >> 5), 1u << (key & 31u));
      atomicAdd(blocksum + (key >> 10), 1u);
    }
  }
}

struct RankScanSmem {
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

// blockpre[b] = sum of blocksum[0..b), blockpre[nb] = number of live seeds
__global__ void __launch_bounds__(kBlock)
rank_scan_kernel(const uint32_t *__restrict__ blocksum, uint32_t *__restrict__ blockpre, uint32_t nb, ChainWs *ws) {
  __shared__ RankScanSmem sm;
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(nb, p, gridDim.x, kBlock, &begin, &end);
  unsigned long long partial = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock) partial += __ldg(blocksum + i);
  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) blockpre[nb] = (uint32_t)(base + chunk_total);
  for (uint32_t t0 = begin; t0 < end; t0 += kBlock) {
    const uint32_t i = t0 + threadIdx.x;
    const uint32_t v = i < end ? __ldg(blocksum + i) : 0u;
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(v, sm.warp, &tile_total);
    if (i < end) blockpre[i] = (uint32_t)base + excl;
    base += tile_total;
  }
  chain_finish(ws, &sm.chain);
}

// keys_out / order in ascending seed id; rows that produce nothing go behind the live ones (their order among
// themselves is irrelevant: EMPTY keys draw and emit nothing)
__global__ void __launch_bounds__(kBlock)
rank_place_kernel(const uint32_t *__restrict__ keys_in, uint32_t n_max, const uint32_t *__restrict__ bitmap,
                  const uint32_t *__restrict__ blockpre, uint32_t nb, uint32_t *counters,
                  uint32_t *__restrict__ keys_out, uint32_t *__restrict__ order) {
  const uint32_t live_total = __ldg(blockpre + nb);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n_max; i += gridDim.x * kBlock) {
    const uint32_t v = __ldg(keys_in + i);
    uint32_t r;
    if (v != kEmpty) {
      const uint32_t w = v >> 5, w0 = w & ~31u;
      r = __ldg(blockpre + (v >> 10));
      for (uint32_t q = w0; q < w; ++q) r += __popc(__ldg(bitmap + q));
      r += __popc(__ldg(bitmap + w) & ((1u << (v & 31u)) - 1u));
    } else {
      r = live_total + atomicAdd(counters, 1u);
    }
    keys_out[r] = v;
    order[r] = i;
  }
}

// restore the all-zero invariant of the rank workspace
__global__ void __launch_bounds__(kBlock)
rank_clear_kernel(const uint32_t *__restrict__ keys_in, uint32_t n_max, uint32_t *bitmap, uint32_t *blocksum,
                  uint32_t *counters) {
  if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = 0u;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n_max; i += gridDim.x * kBlock) {
    const uint32_t v = __ldg(keys_in + i);
    if (v != kEmpty) {
      bitmap[v >> 5] = 0u;
      blocksum[v >> 10] = 0u;
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(kBlock)
replace_sample_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                      const float *__restrict__ prob, const uint32_t *__restrict__ alias,
                      const float *__restrict__ prefix, const uint32_t *__restrict__ keys,
                      const uint32_t *__restrict__ order, uint32_t n_max, uint32_t fanout, RngKey key,
                      uint32_t *__restrict__ tmp_dst) {
  const uint64_t total = (uint64_t)n_max * fanout;
  const uint64_t stride = (uint64_t)gridDim.x * kBlock;
  for (uint64_t e = (uint64_t)blockIdx.x * kBlock + threadIdx.x; e < total; e += stride) {
    const uint32_t r = (uint32_t)(e / fanout);
    const uint32_t j = (uint32_t)(e - (uint64_t)r * fanout);
    const uint32_t v = __ldg(keys + r);
    if (v == kEmpty) continue;  // sorted last; nothing to draw
    const uint32_t i = __ldg(order + r);
    const uint32_t off = __ldg(indptr + v);
    const uint32_t len = __ldg(indptr + v + 1) - off;
    uint32_t dst;
    if (KIND == 1) {  // cuda_sampling_khop1.cu:64-68
      const uint32_t k = rand_u32(key, i, j) % len;
      dst = __ldg(indices + (size_t)off + k);
    } else if (KIND == 2) {  // cuda_sampling_weighted_khop.cu:63-72
      const uint4 b = philox_block(key, i, j >> 1);  // draws 2j, 2j+1 share a block
      const uint32_t r0 = (j & 1) ? b.z : b.x, r1 = (j & 1) ? b.w : b.y;
      const uint32_t k = r0 % len;
      const float u = uniform_f32(r1);
      dst = (u < __ldg(prob + (size_t)off + k)) ? __ldg(indices + (size_t)off + k)
                                                 : __ldg(alias + (size_t)off + k);
    } else {  // cuda_sampling_weighted_khop_prefix.cu:59,64-87
      const float up = __ldg(prefix + (size_t)off + len - 1);
      const float x = uniform_f32(rand_u32(key, i, j)) * up;
      if (x <= __ldg(prefix + off)) {
        dst = __ldg(indices + off);
      } else {
        size_t lo = off, hi = (size_t)off + len - 1;
        while (hi - lo >= 2) {
          const size_t mid = (lo + hi) >> 1;
          if (__ldg(prefix + mid) >= x) hi = mid; else lo = mid;
        }
        dst = __ldg(indices + hi);
      }
    }
    tmp_dst[e] = dst;
  }
}

struct CompactSmem {
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

// keep entry e=(r,j) iff its row is live and (src,dst) differs from entry e+1
// (cuda_sampling_khop1.cu:74-127)
__device__ __forceinline__ uint32_t keep_flag(const uint32_t *keys, const uint32_t *tmp_dst,
                                              uint64_t e, uint64_t total, uint32_t fanout,
                                              uint32_t *src_out) {
  const uint32_t r = (uint32_t)(e / fanout);
  const uint32_t s = __ldg(keys + r);
  *src_out = s;
  if (s == kEmpty) return 0;
  if (e + 1 >= total) return 1;
  const uint32_t r2 = (uint32_t)((e + 1) / fanout);
  const uint32_t s2 = (r2 == r) ? s : __ldg(keys + r2);
  if (s2 != s) return 1;
  return tmp_dst[e] != tmp_dst[e + 1] ? 1u : 0u;
}

__global__ void __launch_bounds__(kBlock)
replace_compact_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ order,
                       const uint32_t *__restrict__ tmp_dst, uint32_t n_max, uint32_t fanout,
                       uint32_t *__restrict__ out_src, uint32_t *__restrict__ out_dst,
                       uint32_t *__restrict__ out_src_local, uint32_t *d_num_out, ChainWs *ws) {
  __shared__ CompactSmem sm;
  const uint64_t total = (uint64_t)n_max * fanout;
  const uint32_t p = chain_ticket(ws, &sm.chain);
  // chunk over flattened entries (total < 2^32 checked on the host)
  uint32_t begin, end;
  chunk_range((uint32_t)total, p, gridDim.x, kBlock, &begin, &end);

  unsigned long long partial = 0;
  uint32_t s;
  for (uint32_t e = begin + threadIdx.x; e < end; e += kBlock)
    partial += keep_flag(keys, tmp_dst, e, total, fanout, &s);
  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);

  for (uint32_t t0 = begin; t0 < end; t0 += kBlock) {
    const uint32_t e = t0 + threadIdx.x;
    uint32_t flag = 0;
    if (e < end) flag = keep_flag(keys, tmp_dst, e, total, fanout, &s);
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(flag, sm.warp, &tile_total);
    if (flag) {
      const size_t o = (size_t)base + excl;
      out_dst[o] = tmp_dst[e];
      if (out_src) out_src[o] = s;
      if (out_src_local) out_src_local[o] = __ldg(order + e / fanout);
    }
    base += tile_total;
  }
  chain_finish(ws, &sm.chain);
}

// ---------------------------------------------------------------------------
// weighted sampling without replacement (alias + rejection)
// ---------------------------------------------------------------------------
constexpr int kHdTile = 64;  // seeds per tile (warp-per-seed selection with memory reads)

struct HdSmem {
  uint32_t rid[kHdTile];
  uint32_t off[kHdTile];
  uint32_t deg[kHdTile];
  uint32_t out[kHdTile + 1];
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

__global__ void __launch_bounds__(kBlock)
hash_dedup_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                  const float *__restrict__ prob, const uint32_t *__restrict__ alias,
                  const uint32_t *__restrict__ input, uint32_t n_max, const uint32_t *__restrict__ d_n,
                  uint32_t fanout, RngKey key, uint32_t *__restrict__ out_src,
                  uint32_t *__restrict__ out_dst, uint32_t *__restrict__ out_src_local,
                  uint32_t *d_num_out, ChainWs *ws) {
  extern __shared__ uint32_t s_acc[];  // [kHdTile][fanout] accepted ids
  __shared__ HdSmem sm;
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, kHdTile, &begin, &end);

  unsigned long long partial = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock) {
    const uint32_t v = __ldg(input + i);
    const uint32_t deg = __ldg(indptr + v + 1) - __ldg(indptr + v);
    partial += deg < fanout ? deg : fanout;
  }
  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t t0 = begin; t0 < end; t0 += kHdTile) {
    uint32_t cnt = 0;
    if (threadIdx.x < kHdTile) {
      const uint32_t i = t0 + threadIdx.x;
      uint32_t deg = 0;
      if (i < end) {
        const uint32_t v = __ldg(input + i);
        const uint32_t o = __ldg(indptr + v);
        deg = __ldg(indptr + v + 1) - o;
        sm.rid[threadIdx.x] = v;
        sm.off[threadIdx.x] = o;
        cnt = deg < fanout ? deg : fanout;
      }
      sm.deg[threadIdx.x] = deg;
    }
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(cnt, sm.warp, &tile_total);
    if (threadIdx.x < kHdTile) sm.out[threadIdx.x] = excl;
    if (threadIdx.x == 0) sm.out[kHdTile] = tile_total;
    __syncthreads();

    for (int s = warp; s < kHdTile; s += kBlock / 32) {
      const uint32_t deg = sm.deg[s];
      if (deg <= fanout) continue;
      const uint32_t off = sm.off[s], item = t0 + s;
      uint32_t *acc = s_acc + s * fanout;
      uint32_t got = 0;
      for (uint32_t tb = 0; got < fanout; tb += 32) {  // hash_dedup.cu:97-111
        const uint32_t t = tb + lane;
        const uint4 b = philox_block(key, item, t >> 1);  // draws 2t, 2t+1
        const uint32_t r0 = (t & 1) ? b.z : b.x, r1 = (t & 1) ? b.w : b.y;
        const uint32_t k = r0 % deg;
        const float u = uniform_f32(r1);
        uint32_t cand = __ldg(indices + (size_t)off + k);
        if (u > __ldg(prob + (size_t)off + k)) cand = __ldg(alias + (size_t)off + k);
        bool fresh = true;
        if (t < FGNN_HASH_DEDUP_MAX_DRAWS) {
          for (uint32_t q = 0; q < got; ++q) fresh = fresh && (acc[q] != cand);
          const uint32_t same = __match_any_sync(0xFFFFFFFFu, cand);
          fresh = fresh && ((__ffs(same) - 1) == lane);
        }
        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, fresh);
        const uint32_t rank = __popc(fm & ((1u << lane) - 1u));
        if (fresh && got + rank < fanout) acc[got + rank] = cand;
        const uint32_t add = __popc(fm);
        got = (got + add < fanout) ? got + add : fanout;
        __syncwarp();
      }
    }
    __syncthreads();

    for (uint32_t e = threadIdx.x; e < tile_total; e += kBlock) {
      uint32_t lo = 0, hi = kHdTile;
#pragma unroll
      for (int it = 0; it < 6; ++it) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sm.out[mid] <= e) lo = mid; else hi = mid;
      }
      const uint32_t s = lo, j = e - sm.out[s];
      const uint32_t nbr = (sm.deg[s] > fanout) ? s_acc[s * fanout + j]
                                                : __ldg(indices + (size_t)sm.off[s] + j);
      const size_t o = (size_t)base + e;
      out_dst[o] = nbr;
      if (out_src) out_src[o] = sm.rid[s];
      if (out_src_local) out_src_local[o] = t0 + s;
    }
    base += tile_total;
    __syncthreads();
  }
  chain_finish(ws, &sm.chain);
}

}  // namespace
}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_sample_replace_workspace_bytes(uint32_t n_max, uint32_t fanout) {
  ReplaceWs w;
  return carve(&w, nullptr, n_max ? n_max : 1, fanout ? fanout : 1);
}

// rank workspace for graphs of up to 2^29 vertices (64 MB bitmap); larger id spaces use the CUB sort
extern "C" size_t fgnn_k_seed_rank_workspace_bytes(size_t num_nodes) {
  if (num_nodes == 0 || num_nodes > ((size_t)1 << 29)) return 0;
  RankWs w;
  return carve_rank(&w, nullptr, num_nodes);
}

extern "C" int fgnn_k_sample_replace(int kind, const uint32_t *indptr, const uint32_t *indices,
                                     const float *prob_table, const uint32_t *alias_table,
                                     const float *prob_prefix_table, const uint32_t *input,
                                     uint32_t n_max, const uint32_t *d_n, uint32_t fanout, fgnn_rng rng,
                                     uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
                                     uint32_t *d_num_out, void *workspace, size_t workspace_bytes,
                                     void *chain_ws, fgnn_stream_t stream) {
  return fgnn_k_sample_replace_ranked(kind, indptr, indices, prob_table, alias_table, prob_prefix_table, input, n_max,
                                      d_n, fanout, rng, out_src, out_dst, out_src_local, d_num_out, workspace,
                                      workspace_bytes, chain_ws, nullptr, 0, stream);
}

extern "C" int fgnn_k_sample_replace_ranked(int kind, const uint32_t *indptr, const uint32_t *indices,
                                            const float *prob_table, const uint32_t *alias_table,
                                            const float *prob_prefix_table, const uint32_t *input,
                                            uint32_t n_max, const uint32_t *d_n, uint32_t fanout, fgnn_rng rng,
                                            uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
                                            uint32_t *d_num_out, void *workspace, size_t workspace_bytes,
                                            void *chain_ws, void *rank_ws, size_t num_nodes,
                                            fgnn_stream_t stream) {
  if (!indptr || !indices || !out_dst || !d_num_out || !chain_ws) return FGNN_ERR_BAD_ARG;
  if (kind != 1 && kind != 2 && kind != 4) return FGNN_ERR_BAD_ARG;
  if (kind == 2 && (!prob_table || !alias_table)) return FGNN_ERR_BAD_ARG;
  if (kind == 4 && !prob_prefix_table) return FGNN_ERR_BAD_ARG;
  if (fanout == 0) return FGNN_ERR_UNSUPPORTED;
  if ((uint64_t)n_max * fanout >= 0xFFFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (n_max == 0) return (int)cudaMemsetAsync(d_num_out, 0, sizeof(uint32_t), st);
  if (!input || !workspace) return FGNN_ERR_BAD_ARG;
  if (workspace_bytes < fgnn_k_sample_replace_workspace_bytes(n_max, fanout)) return FGNN_ERR_BAD_ARG;
  ReplaceWs w;
  carve(&w, workspace, n_max, fanout);
  const RngKey key = make_rng_key(rng);

  const bool ranked = rank_ws != nullptr && fgnn_k_seed_rank_workspace_bytes(num_nodes) != 0;
  if (ranked) {
    RankWs rw;
    carve_rank(&rw, rank_ws, num_nodes);
    const int g1 = persistent_grid(n_max, kBlock, 8, false);
    replace_keys_kernel<<<g1, kBlock, 0, st>>>(indptr, input, n_max, d_n, w.keys_in, w.vals_in, rw.bitmap, rw.blocksum);
    rank_scan_kernel<<<persistent_grid(rw.nb, kBlock * 8, 4, true), kBlock, 0, st>>>(rw.blocksum, rw.blockpre, rw.nb,
                                                                                    (ChainWs *)chain_ws);
    rank_place_kernel<<<g1, kBlock, 0, st>>>(w.keys_in, n_max, rw.bitmap, rw.blockpre, rw.nb, rw.counters, w.keys_out,
                                             w.vals_out);
    rank_clear_kernel<<<g1, kBlock, 0, st>>>(w.keys_in, n_max, rw.bitmap, rw.blocksum, rw.counters);
    note_launch(3);
  } else {
    replace_keys_kernel<<<persistent_grid(n_max, kBlock, 8, false), kBlock, 0, st>>>(
        indptr, input, n_max, d_n, w.keys_in, w.vals_in, nullptr, nullptr);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(w.cub_temp, w.cub_bytes, w.keys_in, w.keys_out,
                                                    w.vals_in, w.vals_out, (int64_t)n_max, 0, 32, st);
    if (e != cudaSuccess) return (int)e;
  }
  const uint64_t total = (uint64_t)n_max * fanout;
  const int g = persistent_grid(total, kBlock, 8, false);
  if (kind == 1)
    replace_sample_kernel<1><<<g, kBlock, 0, st>>>(indptr, indices, prob_table, alias_table,
                                                   prob_prefix_table, w.keys_out, w.vals_out, n_max,
                                                   fanout, key, w.tmp_dst);
  else if (kind == 2)
    replace_sample_kernel<2><<<g, kBlock, 0, st>>>(indptr, indices, prob_table, alias_table,
                                                   prob_prefix_table, w.keys_out, w.vals_out, n_max,
                                                   fanout, key, w.tmp_dst);
  else
    replace_sample_kernel<4><<<g, kBlock, 0, st>>>(indptr, indices, prob_table, alias_table,
                                                   prob_prefix_table, w.keys_out, w.vals_out, n_max,
                                                   fanout, key, w.tmp_dst);
  replace_compact_kernel<<<persistent_grid(total, kBlock, 8, true), kBlock, 0, st>>>(
      w.keys_out, w.vals_out, w.tmp_dst, n_max, fanout, out_src, out_dst, out_src_local, d_num_out,
      (ChainWs *)chain_ws);
  note_launch(3);
  return check_last();
}

extern "C" int fgnn_k_sample_weighted_hash_dedup(
    const uint32_t *indptr, const uint32_t *indices, const float *prob_table,
    const uint32_t *alias_table, const uint32_t *input, uint32_t n_max, const uint32_t *d_n,
    uint32_t fanout, fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
    uint32_t *d_num_out, void *chain_ws, fgnn_stream_t stream) {
  if (!indptr || !indices || !prob_table || !alias_table || !out_dst || !d_num_out || !chain_ws)
    return FGNN_ERR_BAD_ARG;
  if (n_max > 0 && !input) return FGNN_ERR_BAD_ARG;
  if (fanout == 0 || fanout > 128) return FGNN_ERR_UNSUPPORTED;
  if ((uint64_t)n_max * fanout > 0xFFFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  const size_t smem = (size_t)kHdTile * fanout * sizeof(uint32_t);
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, hash_dedup_kernel, kBlock, smem);
  if (occ < 1) occ = 1;
  const int grid = persistent_grid(n_max, kHdTile, occ, true);
  hash_dedup_kernel<<<grid, kBlock, smem, (cudaStream_t)stream>>>(
      indptr, indices, prob_table, alias_table, input, n_max, d_n, fanout, make_rng_key(rng), out_src,
      out_dst, out_src_local, d_num_out, (ChainWs *)chain_ws);
  note_launch();
  return check_last();
}
