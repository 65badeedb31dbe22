// Uniform k-hop neighbour sampling without replacement, fused with COO
// compaction.  Replaces sample_khop0/sample_khop2 + count_edge + DeviceScan +
// compact_edge (reference: cuda_sampling_khop0.cu:42-253, cuda_sampling_khop2.cu:
// 42-252) with ONE persistent launch:
//
//   phase 1  thread-per-seed: read (indptr[v], indptr[v+1]) -> cnt = min(deg,f)
//            chunk aggregate -> chunk-chained scan (common.cuh) -> global base
//   phase 2  per 256-seed tile: block scan of cnt; warp-per-seed selection of
//            the f positions (Philox, all in registers/shared memory — no
//            memory traffic); then an *edge-parallel* gather: every thread owns
//            output slots, finds its (seed, j) by binary search in shared
//            memory, issues the 4-byte neighbour gather and writes the compact
//            COO coalesced.  All lanes are busy and each thread keeps several
//            independent gathers in flight, which is what the HBM-latency-bound
//            4-byte gathers need.
//
// variant 2 (Fisher-Yates, the default sampler of the training scripts) keeps
// the reference's draw sequence r_j % (len-j) but applies the swaps to a
// virtual copy held across the warp (lane t owns the write made at step t), so
// the CSR in HBM is read-only.  variant 0 (reservoir / Algorithm R) evaluates
// the len-f replacement draws in parallel and resolves "last writer wins" with
// a shared-memory atomicMax.
#include "hashtable.cuh"

namespace fgnn {
namespace {

constexpr int kTile = kBlock;  // seeds per tile

struct TileSmem {
  uint32_t rid[kTile];
  uint32_t off[kTile];
  uint32_t deg[kTile];
  uint32_t out[kTile + 1];  // exclusive edge offsets inside the tile
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

// Fisher-Yates over a virtual array, one warp per seed, NS slots per lane.
template <int NS>
__device__ __forceinline__ void select_fisher_yates(const RngKey &key, uint32_t item,
                                                    uint32_t deg, uint32_t fanout,
                                                    uint32_t *choice) {
  const int lane = threadIdx.x & 31;
  uint32_t kmine[NS], mkey[NS], mval[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const uint32_t j = lane + 32 * s;
    mkey[s] = kEmpty;
    mval[s] = 0;
    kmine[s] = 0;
    if (j < fanout) kmine[s] = rand_u32(key, item, j) % (deg - j);
  }
  for (uint32_t j = 0; j < fanout; ++j) {
    const int owner = j & 31, slot = j >> 5;
    uint32_t k = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (s == slot) k = __shfl_sync(0xFFFFFFFFu, kmine[s], owner);
    const uint32_t last = deg - j - 1;
    uint32_t vk = k, vlast = last;
    bool fk = false, fl = false;
    // latest earlier write wins: scan slots from high to low, lanes from high
#pragma unroll
    for (int s = NS - 1; s >= 0; --s) {
      const uint32_t mk = __ballot_sync(0xFFFFFFFFu, mkey[s] == k);
      const uint32_t ml = __ballot_sync(0xFFFFFFFFu, mkey[s] == last);
      const uint32_t cand_k = __shfl_sync(0xFFFFFFFFu, mval[s], mk ? 31 - __clz(mk) : 0);
      const uint32_t cand_l = __shfl_sync(0xFFFFFFFFu, mval[s], ml ? 31 - __clz(ml) : 0);
      if (!fk && mk) { vk = cand_k; fk = true; }
      if (!fl && ml) { vlast = cand_l; fl = true; }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (s == slot && lane == owner) {
        mkey[s] = k;
        mval[s] = vlast;
        choice[j] = vk;
      }
    }
  }
}

// Algorithm R for j = fanout..deg-1, `nthreads` cooperating threads.
__device__ __forceinline__ void select_reservoir(const RngKey &key, uint32_t item,
                                                 uint32_t deg, uint32_t fanout,
                                                 uint32_t *choice, uint32_t tid,
                                                 uint32_t nthreads) {
  const uint32_t ndraw = deg - fanout;
  // each thread evaluates whole Philox blocks (4 draws)
  for (uint32_t b = tid; b * 4u < ndraw; b += nthreads) {
    const uint4 r = philox_block(key, item, b);
#pragma unroll
    for (uint32_t w = 0; w < 4; ++w) {
      const uint32_t d = b * 4u + w;
      if (d < ndraw) {
        const uint32_t j = fanout + d;
        const uint32_t k = pick_word(r, w) % (j + 1u);
        if (k < fanout) atomicMax(&choice[k], j);
      }
    }
  }
}

constexpr uint32_t kCtaSeedDegree = 4096;  // reservoir: above this a whole CTA helps

template <int VARIANT, int NS>
__global__ void __launch_bounds__(kBlock)
sample_khop_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                   const uint32_t *__restrict__ input, uint32_t n_max,
                   const uint32_t *__restrict__ d_n, uint32_t fanout, RngKey key,
                   uint32_t *__restrict__ out_src, uint32_t *__restrict__ out_dst,
                   uint32_t *__restrict__ out_src_local, uint32_t *__restrict__ d_num_out,
                   ChainWs *ws, HtInsert ht) {
  extern __shared__ uint32_t s_choice[];  // [kTile][fanout]
  __shared__ TileSmem sm;

  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, kTile, &begin, &end);

  // ---- phase 1: chunk aggregate ------------------------------------------
  unsigned long long partial = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += kBlock) {
    const uint32_t v = __ldg(input + i);
    const uint32_t deg = __ldg(indptr + v + 1) - __ldg(indptr + v);
    partial += deg < fanout ? deg : fanout;
  }
  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);

  // ---- phase 2: tiles -------------------------------------------------------
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t t0 = begin; t0 < end; t0 += kTile) {
    const uint32_t i = t0 + threadIdx.x;
    uint32_t cnt = 0;
    if (i < end) {
      const uint32_t v = __ldg(input + i);
      const uint32_t o = __ldg(indptr + v);
      const uint32_t deg = __ldg(indptr + v + 1) - o;
      sm.rid[threadIdx.x] = v;
      sm.off[threadIdx.x] = o;
      sm.deg[threadIdx.x] = deg;
      cnt = deg < fanout ? deg : fanout;
    } else {
      sm.deg[threadIdx.x] = 0;
    }
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(cnt, sm.warp, &tile_total);
    sm.out[threadIdx.x] = excl;
    if (threadIdx.x == kBlock - 1) sm.out[kTile] = tile_total;
    // identity choice (covers deg <= fanout and the reservoir's initial fill)
    for (uint32_t e = threadIdx.x; e < kTile * fanout; e += kBlock) s_choice[e] = e % fanout;
    __syncthreads();

    // selection, warp per seed
    for (int s = warp; s < kTile; s += kBlock / 32) {
      const uint32_t deg = sm.deg[s];
      if (deg <= fanout) continue;
      uint32_t *choice = s_choice + s * fanout;
      if (VARIANT == 2) {
        select_fisher_yates<NS>(key, t0 + s, deg, fanout, choice);
      } else {
        if (deg - fanout <= kCtaSeedDegree)
          select_reservoir(key, t0 + s, deg, fanout, choice, lane, 32);
      }
    }
    if (VARIANT == 0) {
      // hub rows: the whole CTA evaluates the replacement draws
      for (int s = 0; s < kTile; ++s) {
        const uint32_t deg = sm.deg[s];
        if (deg > fanout && deg - fanout > kCtaSeedDegree)
          select_reservoir(key, t0 + s, deg, fanout, s_choice + s * fanout, threadIdx.x, kBlock);
      }
    }
    __syncthreads();

    // edge-parallel gather + coalesced compact write
    for (uint32_t e = threadIdx.x; e < tile_total; e += kBlock) {
      // largest s with out[s] <= e
      uint32_t lo = 0, hi = kTile;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sm.out[mid] <= e) lo = mid; else hi = mid;
      }
      const uint32_t s = lo;
      const uint32_t j = e - sm.out[s];
      const uint32_t k = s_choice[s * fanout + j];
      const uint32_t nbr = __ldg(indices + (size_t)sm.off[s] + k);
      const size_t o = (size_t)base + e;
      out_dst[o] = nbr;
      if (out_src) out_src[o] = sm.rid[s];
      if (out_src_local) out_src_local[o] = t0 + s;
      if (ht.table) {
        const uint32_t hp = hash_id(nbr, ht.mask);
        ht.pos_out[o] = insert_item(ht.table, ht.mask, nbr, (uint32_t)o, hp, load_bucket(ht.table + hp));
      }
    }
    base += tile_total;
    __syncthreads();
  }
  chain_finish(ws, &sm.chain);
}

// ---------------------------------------------------------------------------
// variant 2, production path: thread-per-seed Fisher-Yates.
//
// The warp-cooperative selection above costs a chain of dependent shuffles per
// step and processes a warp's 32 seeds one after another (ncu r1_a: 48-52 us for
// 8k and 100k seeds alike, 12-27 % warps active).  Here every thread runs the
// f-step virtual Fisher-Yates of its own seed against a per-thread (key,value)
// log in shared memory ([step][thread] layout, conflict free): ~f^2/2 LDS pairs
// with no cross-lane dependency, 32 seeds per warp in flight.  The neighbour
// gather stays edge-parallel (coalesced compact writes, 4 loads in flight per
// thread).  Tiles are NT seeds (64 or 128) so small layers still spread over
// the whole chip.
// ---------------------------------------------------------------------------
template <int NT>
struct Tile2Smem {
  uint32_t rid[NT];
  uint32_t off[NT];
  uint32_t deg[NT];
  uint32_t out[NT + 1];
  uint32_t warp[NT / 32 + 1];
  ChainSmem chain;
};

template <int NT>
__global__ void __launch_bounds__(NT)
sample_khop2_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                    const uint32_t *__restrict__ input, uint32_t n_max,
                    const uint32_t *__restrict__ d_n, uint32_t fanout, RngKey key,
                    uint32_t *__restrict__ out_src, uint32_t *__restrict__ out_dst,
                    uint32_t *__restrict__ out_src_local, uint32_t *__restrict__ d_num_out,
                    ChainWs *ws, HtInsert ht) {
  extern __shared__ uint32_t dyn[];
  __shared__ Tile2Smem<NT> sm;
  const uint32_t fs = fanout | 1u;            // odd row stride: conflict-free [seed][j]
  uint32_t *s_choice = dyn;                   // [NT][fs]
  uint32_t *s_mkey = dyn + NT * fs;           // [fanout][NT]
  uint32_t *s_mval = s_mkey + NT * fanout;    // [fanout][NT]

  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, NT, &begin, &end);

  unsigned long long partial = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += NT) {
    const uint32_t v = __ldg(input + i);
    const uint32_t deg = __ldg(indptr + v + 1) - __ldg(indptr + v);
    partial += deg < fanout ? deg : fanout;
  }
  unsigned long long chunk_total;
  unsigned long long base = chain_scan<NT>(ws, &sm.chain, p, partial, &chunk_total);
  if (p == gridDim.x - 1 && threadIdx.x == 0) *d_num_out = (uint32_t)(base + chunk_total);

  const uint32_t tid = threadIdx.x;
  for (uint32_t t0 = begin; t0 < end; t0 += NT) {
    const uint32_t i = t0 + tid;
    uint32_t cnt = 0, deg = 0;
    if (i < end) {
      const uint32_t v = __ldg(input + i);
      const uint32_t o = __ldg(indptr + v);
      deg = __ldg(indptr + v + 1) - o;
      sm.rid[tid] = v;
      sm.off[tid] = o;
      cnt = deg < fanout ? deg : fanout;
    }
    sm.deg[tid] = deg;
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan<NT>(cnt, sm.warp, &tile_total);
    sm.out[tid] = excl;
    if (tid == NT - 1) sm.out[NT] = tile_total;

    uint32_t *choice = s_choice + tid * fs;
    if (deg > fanout) {  // cuda_sampling_khop2.cu:72-83 on a virtual copy of the row
      uint4 blk = make_uint4(0, 0, 0, 0);
      for (uint32_t j = 0; j < fanout; ++j) {
        if ((j & 3u) == 0) blk = philox_block(key, i, j >> 2);
        const uint32_t k = pick_word(blk, j & 3u) % (deg - j);
        const uint32_t last = deg - j - 1;
        uint32_t vk = k, vlast = last;
#pragma unroll 4
        for (uint32_t t = 0; t < j; ++t) {  // latest write wins
          const uint32_t kk = s_mkey[t * NT + tid];
          const uint32_t vv = s_mval[t * NT + tid];
          if (kk == k) vk = vv;
          if (kk == last) vlast = vv;
        }
        s_mkey[j * NT + tid] = k;
        s_mval[j * NT + tid] = vlast;
        choice[j] = vk;
      }
    } else {
      for (uint32_t j = 0; j < cnt; ++j) choice[j] = j;
    }
    __syncthreads();

    for (uint32_t e0 = tid; e0 < tile_total; e0 += NT * 4) {
      uint32_t nbr[4], ss[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t e = e0 + u * NT;
        ss[u] = kEmpty;
        if (e < tile_total) {
          uint32_t lo = 0, hi = NT;
          while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sm.out[mid] <= e) lo = mid; else hi = mid;
          }
          const uint32_t k = s_choice[lo * fs + (e - sm.out[lo])];
          nbr[u] = __ldg(indices + (size_t)sm.off[lo] + k);
          ss[u] = lo;
        }
      }
      // the picked ids go straight into the batch's OrderedHashTable (owner = smallest output index):
      // first probes of the four ids are issued together, behind the gathers that produced them
      uint2 hb[4];
      uint32_t hp[4];
      if (ht.table) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          hp[u] = hash_id(nbr[u], ht.mask);
          if (ss[u] != kEmpty) hb[u] = load_bucket(ht.table + hp[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (ss[u] != kEmpty) {
          const size_t o = (size_t)base + e0 + u * NT;
          out_dst[o] = nbr[u];
          if (out_src) out_src[o] = sm.rid[ss[u]];
          if (out_src_local) out_src_local[o] = t0 + ss[u];
          if (ht.table) ht.pos_out[o] = insert_item(ht.table, ht.mask, nbr[u], (uint32_t)o, hp[u], hb[u]);
        }
      }
    }
    base += tile_total;
    __syncthreads();
  }
  chain_finish<NT>(ws, &sm.chain);
}

template <int NT>
int launch2(const uint32_t *indptr, const uint32_t *indices, const uint32_t *input, uint32_t n_max,
            const uint32_t *d_n, uint32_t fanout, RngKey key, uint32_t *out_src, uint32_t *out_dst,
            uint32_t *out_src_local, uint32_t *d_num_out, void *chain_ws, HtInsert ht, cudaStream_t stream) {
  const size_t smem = ((size_t)NT * (fanout | 1u) + 2 * (size_t)NT * fanout) * sizeof(uint32_t);
  auto kern = sample_khop2_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = persistent_grid(n_max, NT, occupancy(kern, NT, smem), true);
  kern<<<grid, NT, smem, stream>>>(indptr, indices, input, n_max, d_n, fanout, key, out_src, out_dst,
                                   out_src_local, d_num_out, (ChainWs *)chain_ws, ht);
  note_launch();
  return check_last();
}


// ---------------------------------------------------------------------------
// variant 2 inside fgnn_k_sample_batch: PADDED output, no compaction here.
//
// The kernel above spends most of its ~20-40 us on things that are not sampling: a first pass over the
// seeds to count edges, the chunk-chained scan that turns the counts into output offsets (a chain of
// dependent L2 round trips across up to 688 tickets), a block scan and a binary search per edge.  Inside
// a mini-batch none of that is needed: the unique/remap pass that follows already runs ONE chained scan,
// and it can carry the edge offsets next to the new-id offsets.  So here every seed simply owns `fanout`
// output slots: dst[i*fanout + j] = j-th pick or EMPTY.  One thread per seed: three dependent loads
// (id -> indptr pair -> up to `fanout` independent neighbour gathers), the same virtual Fisher-Yates and
// Philox counters as sample_khop2_kernel (so the compacted result is bit-identical), and a coalesced
// write of the tile through shared memory.  No scan, no ticket, no cross-CTA traffic.
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
sample_khop2_pad_kernel(const uint32_t *__restrict__ indptr, const uint32_t *__restrict__ indices,
                        const uint32_t *__restrict__ input, uint32_t n_max,
                        const uint32_t *__restrict__ d_n, uint32_t fanout, RngKey key,
                        uint32_t *__restrict__ out_dst) {
  extern __shared__ uint32_t dyn[];
  const uint32_t fs = fanout | 1u;            // odd row stride: conflict-free [seed][j]
  uint32_t *s_choice = dyn;                   // [NT][fs]   positions, then the gathered ids
  uint32_t *s_mkey = dyn + NT * fs;           // [fanout][NT]
  uint32_t *s_mval = s_mkey + NT * fanout;    // [fanout][NT]
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t tid = threadIdx.x;
  for (uint32_t t0 = blockIdx.x * NT; t0 < n; t0 += gridDim.x * NT) {
    const uint32_t i = t0 + tid;
    uint32_t *choice = s_choice + tid * fs;
    if (i < n) {
      const uint32_t v = __ldg(input + i);
      const uint32_t o = __ldg(indptr + v);
      const uint32_t deg = __ldg(indptr + v + 1) - o;
      uint32_t cnt = deg < fanout ? deg : fanout;
      if (deg > fanout) {  // cuda_sampling_khop2.cu:72-83 on a virtual copy of the row
        uint4 blk = make_uint4(0, 0, 0, 0);
        for (uint32_t j = 0; j < fanout; ++j) {
          if ((j & 3u) == 0) blk = philox_block(key, i, j >> 2);
          const uint32_t k = pick_word(blk, j & 3u) % (deg - j);
          const uint32_t last = deg - j - 1;
          uint32_t vk = k, vlast = last;
#pragma unroll 4
          for (uint32_t t = 0; t < j; ++t) {  // latest write wins
            const uint32_t kk = s_mkey[t * NT + tid];
            const uint32_t vv = s_mval[t * NT + tid];
            if (kk == k) vk = vv;
            if (kk == last) vlast = vv;
          }
          s_mkey[j * NT + tid] = k;
          s_mval[j * NT + tid] = vlast;
          choice[j] = vk;
        }
      } else {
        for (uint32_t j = 0; j < cnt; ++j) choice[j] = j;
      }
      // gather: 8 independent 4-byte loads in flight per thread
      const uint32_t *row = indices + (size_t)o;
      for (uint32_t j0 = 0; j0 < fanout; j0 += 8) {
        uint32_t nb[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t j = j0 + u;
          nb[u] = kEmpty;
          if (j < cnt) nb[u] = __ldg(row + choice[j]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (j0 + u < fanout) choice[j0 + u] = nb[u];
      }
    }
    __syncthreads();
    // the tile's padded block is contiguous: coalesced copy out of shared memory
    const uint32_t rows = n - t0 < (uint32_t)NT ? n - t0 : (uint32_t)NT;
    uint32_t *dst = out_dst + (size_t)t0 * fanout;
    for (uint32_t e = tid; e < rows * fanout; e += NT) {
      const uint32_t r = e / fanout;
      dst[e] = s_choice[r * fs + (e - r * fanout)];
    }
    __syncthreads();
  }
}

template <int NT>
int launch2_pad(const uint32_t *indptr, const uint32_t *indices, const uint32_t *input, uint32_t n_max,
                const uint32_t *d_n, uint32_t fanout, RngKey key, uint32_t *out_dst, cudaStream_t stream) {
  const size_t smem = ((size_t)NT * (fanout | 1u) + 2 * (size_t)NT * fanout) * sizeof(uint32_t);
  auto kern = sample_khop2_pad_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = persistent_grid(n_max, NT, occupancy(kern, NT, smem), false);
  kern<<<grid, NT, smem, stream>>>(indptr, indices, input, n_max, d_n, fanout, key, out_dst);
  note_launch();
  return check_last();
}

template <int VARIANT, int NS>
int launch(const uint32_t *indptr, const uint32_t *indices, const uint32_t *input,
           uint32_t n_max, const uint32_t *d_n, uint32_t fanout, RngKey key,
           uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
           uint32_t *d_num_out, void *chain_ws, HtInsert ht, cudaStream_t stream) {
  const size_t smem = (size_t)kTile * fanout * sizeof(uint32_t);
  auto kern = sample_khop_kernel<VARIANT, NS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, smem);
  if (occ < 1) occ = 1;
  const int grid = persistent_grid(n_max, kTile, occ, true);
  kern<<<grid, kBlock, smem, stream>>>(indptr, indices, input, n_max, d_n, fanout, key, out_src,
                                       out_dst, out_src_local, d_num_out, (ChainWs *)chain_ws, ht);
  note_launch();
  return check_last();
}

}  // namespace

int sample_khop_launch(int variant, const uint32_t *indptr, const uint32_t *indices,
                       const uint32_t *input, uint32_t n_max, const uint32_t *d_n, uint32_t fanout,
                       fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
                       uint32_t *d_num_out, void *chain_ws, HtInsert ht, cudaStream_t st) {
  if (!indptr || !indices || !out_dst || !d_num_out || !chain_ws) return FGNN_ERR_BAD_ARG;
  if (n_max > 0 && !input) return FGNN_ERR_BAD_ARG;
  if (fanout == 0 || fanout > 128) return FGNN_ERR_UNSUPPORTED;
  if ((uint64_t)n_max * fanout > 0x7FFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  if (ht.table && !ht.pos_out) return FGNN_ERR_BAD_ARG;
  const RngKey key = make_rng_key(rng);
#define FGNN_GO(V, NS)                                                                      \
  return launch<V, NS>(indptr, indices, input, n_max, d_n, fanout, key, out_src, out_dst,   \
                       out_src_local, d_num_out, chain_ws, ht, st)
  if (variant == 2) {
    // thread-per-seed Fisher-Yates; 64-seed tiles when the layer is small or the
    // per-thread log would not fit next to a 128-seed tile
    const bool small = (uint64_t)n_max <= (uint64_t)sm_count() * 128ull || fanout > 48;
    if (small)
      return launch2<64>(indptr, indices, input, n_max, d_n, fanout, key, out_src, out_dst,
                         out_src_local, d_num_out, chain_ws, ht, st);
    return launch2<128>(indptr, indices, input, n_max, d_n, fanout, key, out_src, out_dst,
                        out_src_local, d_num_out, chain_ws, ht, st);
  } else if (variant == 22) {  // warp-cooperative Fisher-Yates (kept for A/B profiling)
    if (fanout <= 32) FGNN_GO(2, 1);
    if (fanout <= 64) FGNN_GO(2, 2);
    FGNN_GO(2, 4);
  } else if (variant == 0) {
    FGNN_GO(0, 1);
  }
#undef FGNN_GO
  return FGNN_ERR_BAD_ARG;
}


int sample_khop2_pad_launch(const uint32_t *indptr, const uint32_t *indices, const uint32_t *input,
                            uint32_t n_max, const uint32_t *d_n, uint32_t fanout, fgnn_rng rng,
                            uint32_t *out_dst_padded, cudaStream_t st) {
  if (!indptr || !indices || !out_dst_padded) return FGNN_ERR_BAD_ARG;
  if (n_max > 0 && !input) return FGNN_ERR_BAD_ARG;
  if (fanout == 0 || fanout > 128) return FGNN_ERR_UNSUPPORTED;
  if ((uint64_t)n_max * fanout > 0x7FFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  if (n_max == 0) return 0;
  const RngKey key = make_rng_key(rng);
  const bool small = (uint64_t)n_max <= (uint64_t)sm_count() * 128ull || fanout > 48;
  if (small) return launch2_pad<64>(indptr, indices, input, n_max, d_n, fanout, key, out_dst_padded, st);
  return launch2_pad<128>(indptr, indices, input, n_max, d_n, fanout, key, out_dst_padded, st);
}

}  // namespace fgnn

extern "C" int fgnn_k_sample_khop(int variant, const uint32_t *indptr, const uint32_t *indices,
                                  const uint32_t *input, uint32_t n_max, const uint32_t *d_n,
                                  uint32_t fanout, fgnn_rng rng, uint32_t *out_src,
                                  uint32_t *out_dst, uint32_t *out_src_local,
                                  uint32_t *d_num_out, void *chain_ws, fgnn_stream_t stream) {
  fgnn::HtInsert none{nullptr, 0, nullptr};
  return fgnn::sample_khop_launch(variant, indptr, indices, input, n_max, d_n, fanout, rng, out_src,
                                  out_dst, out_src_local, d_num_out, chain_ws, none,
                                  (cudaStream_t)stream);
}
