// The uniform k-hop mini-batch chain (khop2, the training scripts' default sampler), two launches per layer
// for up to FGNN_MAX_SUPER mini-batches at once ("super-batch": blockIdx.y = mini-batch, every mini-batch has
// its own hash table, unique list, scratch and scan workspace side by side).
//
//   fc_sample_kernel          sample_khop2 (cuda_sampling_khop2.cu:42-90) into a padded [seed][fanout] block
//   fc_insert_kernel          FillWithDuplicates' insert half (cuda_hashtable.cu:49-61,131-174); the first sampled
//                             layer also does FillWithUnique of the seeds (cuda_hashtable.cu:1017-1037)
//   fc_compact_kernel         count_edge + DeviceScan + compact_edge + the numbering half of FillWithDuplicates
//                             + GPUMapEdges (cuda_sampling_khop2.cu:121-175, cuda_hashtable.cu:387-438,725-807,
//                             cuda_mapping.cu:68-81)
//
// What round 1's profile said about the three-kernel chain it replaces (profiles/r1_r_kernels_full.txt):
//   * the sampler was ISSUE-bound: every thread of a warp walked an O(f^2) swap log although only ~1 seed in 12
//     has more than `fanout` neighbours.  Here the seeds that need a Fisher-Yates are compacted into dense lanes
//     first, and each keeps its virtual swaps in a small open-addressed map in shared memory ([slot][thread]
//     layout: the bank depends on the thread only, so random slots never conflict): ~30x fewer instructions.
//   * the neighbour gather is edge-parallel over the padded [seed][fanout] tile (coalesced, 8 loads in flight per
//     thread).  Inserting the picks from inside the sampler was tried twice (round 1 r1_n, round 2 r2_b/r2_c:
//     268-350 us per 4 mini-batches): the sampler's CTAs are phase-structured and shared-memory-limited, too few
//     threads are in the probe at any time to hide the table's latency.  The insert is its own launch again:
//     256-thread CTAs at full occupancy, four independent probes per thread.
//   * compaction: one CTA owns 2048 consecutive padded slots held in registers, counts with warp ballots (two
//     block barriers instead of 24), and the cross-CTA prefix is a block-wide DIRECT sum of the lower tickets'
//     aggregates (one L2 round trip) instead of a warp look-back (~10 dependent round trips for 700 tickets).
//   * the table is never cleared: buckets carry a 7-bit version tag in their local word (reference: 16-byte
//     buckets with a `version` field, cuda_hashtable.cu:714-723); a bucket of another version is free and is
//     claimed with ONE 64-bit compare-and-swap.  The caller clears the table when the tag wraps (every 126
//     batches).
// Results are bit-identical to the three-kernel chain and to oracle.sample_batch_oracle (same Philox counters,
// same first-occurrence ownership rule).
#include "hashtable.cuh"

#include <string.h>

namespace fgnn {
namespace {

constexpr int kFcItems = 8;                     // padded slots per thread in the compaction kernel
constexpr int kFcChunk = kBlock * kFcItems;     // 2048 padded slots per CTA
constexpr uint32_t kFcVerMask = 0x7F000000u;    // version tag inside the bucket's local word

struct FcBatch {
  const uint32_t *seeds;     // first sampled layer: the mini-batch's seeds (unique by contract)
  const uint32_t *d_n_seeds; // optional device count of seeds
  uint32_t *n2o;             // running unique list
  Bucket *table;
  uint32_t *num_items;       // device: unique ids so far
  uint32_t *counts;          // device [L][3] = num_dst, num_edge, num_src
  uint32_t *dst;             // this layer's scratch: sampled global id of every padded slot (EMPTY = hole) ...
  uint32_t *pos;             // ... overwritten in place by the slot's bucket position (same array)
  uint32_t *row, *col;       // this layer's outputs
  ChainWs *ws;
  RngKey key;                // Philox key of (seed, batch_key, layer)
  uint32_t n_seed_max;
  uint32_t vtag;             // version << 24, or 0 when the table was cleared for this batch
};

struct FcArgs {
  FcBatch b[FGNN_MAX_SUPER];
  const uint32_t *indptr, *indices;
  uint32_t mask;             // capacity - 1
  uint32_t vmask;            // kFcVerMask, or 0 (unversioned: emptiness = EMPTY key only)
  uint32_t valmask;          // bits of an assigned local id
  uint32_t layer, first;     // first = 1: inputs are the seeds, which are inserted here too
  uint32_t fanout, n_max;    // n_max bounds the layer's inputs
  uint32_t nfy, hslots, hshift;  // Fisher-Yates lanes per round, map slots per lane (power of two), 32 - log2
};

__device__ __forceinline__ unsigned long long pack_bucket(uint32_t key, uint32_t local) {
  return (unsigned long long)key | ((unsigned long long)local << 32);
}

__device__ __forceinline__ uint2 ld_bucket_cg(const Bucket *b) {
  return __ldcg(reinterpret_cast<const uint2 *>(b));
}

// Insert `id` with candidate local word `mine` (tag | index for a seed, tag | PENDING | padded index for a
// pick): the smallest word wins, so seeds (assigned, < PENDING) beat picks and the smallest padded index owns
// a new id.  A bucket whose tag is not the batch's is free.  Returns the bucket position.
__device__ __forceinline__ uint32_t fc_put(Bucket *table, uint32_t mask, uint32_t vmask, uint32_t vtag,
                                           uint32_t id, uint32_t mine, uint32_t pos, uint2 b) {
  while (true) {
    const bool free_b = (b.x == kEmpty) || ((b.y & vmask) != vtag);
    if (!free_b && b.x == id) {
      if (b.y > mine) atomicMin(&table[pos].local, mine);
      return pos;
    }
    if (free_b) {
      const unsigned long long seen = pack_bucket(b.x, b.y);
      const unsigned long long old =
          atomicCAS(reinterpret_cast<unsigned long long *>(table + pos), seen, pack_bucket(id, mine));
      if (old == seen) return pos;
      b.x = (uint32_t)old;           // somebody claimed it first: look at what is there now
      b.y = (uint32_t)(old >> 32);
      continue;
    }
    pos = (pos + 1) & mask;
    b = ld_bucket_cg(table + pos);
  }
}

// ---------------------------------------------------------------------------------------------------------
// sample + insert
// ---------------------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
fc_sample_kernel(const __grid_constant__ FcArgs a) {
  extern __shared__ __align__(16) uint32_t dyn[];
  const FcBatch &B = a.b[blockIdx.y];
  const uint32_t f = a.fanout, fs = f | 1u;
  uint32_t *s_off = dyn;                    // [NT]
  uint32_t *s_deg = s_off + NT;             // [NT]
  uint32_t *s_list = s_deg + NT;            // [NT] tile-local seeds that need a Fisher-Yates
  uint32_t *s_choice = s_list + NT;         // [NT][fs] picked positions of those seeds
  uint32_t *s_keys = s_choice + ((NT * fs + 3u) & ~3u);  // [hslots][nfy]
  uint32_t *s_vals = s_keys + a.hslots * a.nfy;           // [hslots][nfy]
  __shared__ uint32_t s_nfy;

  const uint32_t tid = threadIdx.x;
  uint32_t n;
  if (a.first) {
    n = load_count(B.n_seed_max, B.d_n_seeds);
    if (blockIdx.x == 0 && tid == 0) {  // FillWithUnique bookkeeping: the seeds are local ids [0, n)
      *B.num_items = n;
      B.counts[3 * a.layer] = n;
    }
  } else {
    n = load_count(a.n_max, B.counts + 3 * a.layer);
  }
  const uint32_t *input = a.first ? B.seeds : B.n2o;
  const uint32_t H = a.hslots, NFY = a.nfy;

  for (uint32_t t0 = blockIdx.x * NT; t0 < n; t0 += gridDim.x * NT) {
    // ---- phase A: seed -> (row offset, degree); seeds of the batch enter the table as assigned ids ----
    const uint32_t i = t0 + tid;
    uint32_t deg = 0, off = 0;
    if (tid == 0) s_nfy = 0;
    if (i < n) {
      const uint32_t v = __ldg(input + i);
      off = __ldg(a.indptr + v);
      deg = __ldg(a.indptr + v + 1) - off;
      if (a.first) B.n2o[i] = v;  // the seeds open the unique list (local id = position)
    }
    s_off[tid] = off;
    s_deg[tid] = deg;
    __syncthreads();
    if (deg > f) s_list[atomicAdd(&s_nfy, 1u)] = tid;
    __syncthreads();
    const uint32_t nfy = s_nfy;

    // ---- phase B: Fisher-Yates of the rows longer than the fanout, dense lanes ---------------------------
    // cuda_sampling_khop2.cu:72-83 on a VIRTUAL copy of the row: position p holds map[p] if present, else p.
    // Step j: k = r_j % (deg-j); pick value(k); position k <- value(deg-j-1).
    for (uint32_t r0 = 0; r0 < nfy; r0 += NFY) {
      for (uint32_t w = tid * 4; w < H * NFY; w += NT * 4)
        *reinterpret_cast<uint4 *>(s_keys + w) = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
      __syncthreads();
      if (tid < NFY && r0 + tid < nfy) {
        const uint32_t s = s_list[r0 + tid];
        const uint32_t d = s_deg[s];
        uint32_t *choice = s_choice + s * fs;
        uint32_t *keys = s_keys + tid, *vals = s_vals + tid;
        uint4 blk = make_uint4(0, 0, 0, 0);
        for (uint32_t j = 0; j < f; ++j) {
          if ((j & 3u) == 0) blk = philox_block(B.key, t0 + s, j >> 2);
          const uint32_t m = d - j;
          const uint32_t k = pick_word(blk, j & 3u) % m;
          const uint32_t last = m - 1;
          uint32_t hk = (k * 0x9E3779B1u) >> a.hshift;
          uint32_t vk = k;
          while (true) {
            const uint32_t kk = keys[hk * NFY];
            if (kk == k) { vk = vals[hk * NFY]; break; }
            if (kk == kEmpty) break;
            hk = (hk + 1) & (H - 1);
          }
          uint32_t vlast = vk;
          if (last != k) {
            vlast = last;
            uint32_t hl = (last * 0x9E3779B1u) >> a.hshift;
            while (true) {
              const uint32_t kk = keys[hl * NFY];
              if (kk == last) { vlast = vals[hl * NFY]; break; }
              if (kk == kEmpty) break;
              hl = (hl + 1) & (H - 1);
            }
          }
          keys[hk * NFY] = k;
          vals[hk * NFY] = vlast;
          choice[j] = vk;
        }
      }
      __syncthreads();
    }

    // ---- phase C: edge-parallel gather of the padded tile, coalesced write (EMPTY = hole) -----------------
    const uint32_t rows = n - t0 < (uint32_t)NT ? n - t0 : (uint32_t)NT;
    const uint32_t tile_items = rows * f;
    uint32_t *dst_out = B.dst + (size_t)t0 * f;
    constexpr int kIlp = 8;
    for (uint32_t e0 = tid; e0 < tile_items; e0 += NT * kIlp) {
      uint32_t nbr[kIlp];
#pragma unroll
      for (int u = 0; u < kIlp; ++u) {
        const uint32_t e = e0 + u * NT;
        nbr[u] = kEmpty;
        if (e < tile_items) {
          const uint32_t s = e / f, j = e - s * f;
          const uint32_t d = s_deg[s];
          if (j < (d < f ? d : f)) {
            const uint32_t p = d > f ? s_choice[s * fs + j] : j;
            nbr[u] = __ldg(a.indices + (size_t)s_off[s] + p);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kIlp; ++u)
        if (e0 + u * NT < tile_items) dst_out[e0 + u * NT] = nbr[u];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// insert: seeds (first layer) as assigned ids, every pick as a candidate owner of its id
// ---------------------------------------------------------------------------------------------------------
constexpr int kFcInsIlp = 4;

__global__ void __launch_bounds__(kBlock)
fc_insert_kernel(const __grid_constant__ FcArgs a) {
  const FcBatch &B = a.b[blockIdx.y];
  const uint32_t n_in = a.first ? load_count(B.n_seed_max, B.d_n_seeds) : load_count(a.n_max, B.counts + 3 * a.layer);
  const uint32_t n_pad = n_in * a.fanout;
  const uint32_t n_seed = a.first ? n_in : 0u;   // seed items come first in the item space
  const uint32_t total = n_seed + n_pad;
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i0 = blockIdx.x * kBlock + threadIdx.x; i0 < total; i0 += stride * kFcInsIlp) {
    uint32_t id[kFcInsIlp], hp[kFcInsIlp], mine[kFcInsIlp];
    uint2 hb[kFcInsIlp];
#pragma unroll
    for (int u = 0; u < kFcInsIlp; ++u) {
      const uint32_t i = i0 + u * stride;
      id[u] = kEmpty;
      mine[u] = 0;
      if (i < n_seed) {
        id[u] = __ldg(B.seeds + i);
        mine[u] = B.vtag | i;
      } else if (i < total) {
        id[u] = __ldcs(B.dst + (i - n_seed));
        mine[u] = B.vtag | kPending | (i - n_seed);
      }
      hp[u] = hash_id(id[u], a.mask);
    }
#pragma unroll
    for (int u = 0; u < kFcInsIlp; ++u) {
      hb[u] = make_uint2(0u, 0u);
      if (id[u] != kEmpty) hb[u] = ld_bucket_cg(B.table + hp[u]);
    }
#pragma unroll
    for (int u = 0; u < kFcInsIlp; ++u) {
      const uint32_t i = i0 + u * stride;
      if (i >= total) continue;
      uint32_t bp = kEmpty;
      if (id[u] != kEmpty) bp = fc_put(B.table, a.mask, a.vmask, B.vtag, id[u], mine[u], hp[u], hb[u]);
      if (i >= n_seed) B.pos[i - n_seed] = bp;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// compact + number the new ids + remap
// ---------------------------------------------------------------------------------------------------------
struct FcCompactSmem {
  uint32_t cnt[kFcItems * (kBlock / 32)];   // per (item round, warp): valid << 16 | new
  uint32_t pre[kFcItems * (kBlock / 32)];   // exclusive prefix of the above
  uint32_t total;                           // chunk total, same packing
  uint32_t ticket, last;
  unsigned long long excl;
  unsigned long long warp64[kBlock / 32];
};

__global__ void __launch_bounds__(kBlock)
fc_compact_kernel(const __grid_constant__ FcArgs a) {
  __shared__ FcCompactSmem sm;
  const FcBatch &B = a.b[blockIdx.y];
  ChainWs *ws = B.ws;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t f = a.fanout;
  const uint32_t n_seed = load_count(a.n_max, B.counts + 3 * a.layer);
  const uint32_t n = n_seed * f;                       // padded slots of this layer
  const uint32_t P = gridDim.x;
  const uint32_t P_n = n ? (n + kFcChunk - 1) / kFcChunk : 1u;  // tickets that own slots (ticket 0 always reports)
  if (tid == 0) sm.ticket = atomicAdd(&ws->ticket, 1u);
  __syncthreads();
  const uint32_t p = sm.ticket;
  const uint32_t items0 = *B.num_items;  // stable: only the CTA that finishes last updates it, at the end
  constexpr unsigned long long kFlag = 1ull << 62, kVal = (1ull << 62) - 1;

  if (p < P_n) {
    const uint32_t begin = p * kFcChunk;
    // ---- load: bucket position and bucket of every slot of the chunk, all independent ----
    uint32_t bp[kFcItems], key[kFcItems], w[kFcItems];
#pragma unroll
    for (int it = 0; it < kFcItems; ++it) {
      const uint32_t i = begin + it * kBlock + tid;
      bp[it] = i < n ? __ldcs(B.pos + i) : kEmpty;
    }
#pragma unroll
    for (int it = 0; it < kFcItems; ++it) {
      key[it] = 0;
      w[it] = 0;
      if (bp[it] != kEmpty) {
        const uint2 b = ld_bucket_cg(B.table + bp[it]);
        key[it] = b.x;
        w[it] = b.y;
      }
    }
    // ---- count with ballots: one smem word per (round, warp) ----
    uint32_t vm[kFcItems], nm[kFcItems];
#pragma unroll
    for (int it = 0; it < kFcItems; ++it) {
      const uint32_t i = begin + it * kBlock + tid;
      const bool valid = bp[it] != kEmpty;
      const bool isnew = valid && w[it] == (B.vtag | kPending | i);
      vm[it] = __ballot_sync(0xFFFFFFFFu, valid);
      nm[it] = __ballot_sync(0xFFFFFFFFu, isnew);
      if (lane == 0) sm.cnt[it * (kBlock / 32) + warp] = ((uint32_t)__popc(vm[it]) << 16) | (uint32_t)__popc(nm[it]);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the 64 (round, warp) counts: two per lane
      const uint32_t c0 = sm.cnt[2 * lane], c1 = sm.cnt[2 * lane + 1];
      const uint32_t incl = warp_incl_scan(c0 + c1);
      sm.pre[2 * lane] = incl - c0 - c1;
      sm.pre[2 * lane + 1] = incl - c1;
      if (lane == 31) {
        sm.total = incl;
        // publish this chunk's aggregate {edges : 31 | new ids : 31}
        st_relaxed_u64(&ws->agg[p], kFlag | ((unsigned long long)(incl >> 16) << 31) | (incl & 0xFFFFu));
      }
    }
    // ---- cross-CTA prefix: block-wide direct sum of the lower tickets' aggregates ----
    unsigned long long c = 0;
    for (uint32_t t = tid; t < p; t += kBlock) {
      unsigned long long v = ld_relaxed_u64(&ws->agg[t]);
      while (!(v >> 62)) {
        __nanosleep(40);
        v = ld_relaxed_u64(&ws->agg[t]);
      }
      c += v & kVal;
    }
    const unsigned long long excl = block_sum_u64(c, sm.warp64);  // contains the barriers that publish sm.pre
    const uint32_t base_edge = (uint32_t)(excl >> 31), base_new = (uint32_t)(excl & 0x7FFFFFFFull);
    const uint32_t lt = (1u << lane) - 1u;

    // ---- owners: local id, unique list entry, and their own edge ----
    uint32_t eo[kFcItems];
#pragma unroll
    for (int it = 0; it < kFcItems; ++it) {
      const uint32_t i = begin + it * kBlock + tid;
      const uint32_t pre = sm.pre[it * (kBlock / 32) + warp];
      eo[it] = base_edge + (pre >> 16) + (uint32_t)__popc(vm[it] & lt);
      if (bp[it] != kEmpty) {
        B.col[eo[it]] = i / f;  // the seed's local id: layer inputs are the first entries of the unique list
        if (nm[it] & (1u << lane)) {
          const uint32_t local = items0 + base_new + (pre & 0xFFFFu) + (uint32_t)__popc(nm[it] & lt);
          B.table[bp[it]].local = B.vtag | local;
          B.n2o[local] = key[it];
          B.row[eo[it]] = local;
        }
      }
    }
    __syncthreads();  // this CTA's owners are assigned and visible to the CTA
    // ---- everybody else reads the owner's local id; only lower tickets can still be pending ----
#pragma unroll
    for (int it = 0; it < kFcItems; ++it) {
      if (bp[it] != kEmpty && !(nm[it] & (1u << lane))) {
        uint32_t v = w[it];
        if (v & kPending) v = wait_local_word(B.table, bp[it]);
        B.row[eo[it]] = v & a.valmask;
      }
    }
    if (p == P_n - 1 && tid == 0) {  // totals of the layer travel through the pad words
      const uint32_t tot = sm.total;
      ws->pad[0] = base_new + (tot & 0xFFFFu);
      ws->pad[1] = base_edge + (tot >> 16);
    }
  }

  // the CTA that finishes LAST publishes the counts (every other CTA has read items0 by then) and re-arms
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    sm.last = (prev == P - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sm.last) {
    __threadfence();
    for (uint32_t t = tid; t < P && t < (uint32_t)kMaxChainCtas; t += kBlock) ws->agg[t] = 0ull;
    if (tid == 0) {
      const uint32_t total = items0 + *((volatile unsigned int *)&ws->pad[0]);
      const uint32_t edges = *((volatile unsigned int *)&ws->pad[1]);
      *B.num_items = total;
      B.counts[3 * a.layer + 1] = edges;
      B.counts[3 * a.layer + 2] = total;
      if (a.layer > 0) B.counts[3 * (a.layer - 1)] = total;
      ws->pad[0] = 0u;
      ws->pad[1] = 0u;
      ws->ticket = 0u;
      ws->done = 0u;
    }
  }
}

template <int NT>
int launch_sample(const FcArgs &a, uint32_t K, cudaStream_t st) {
  const uint32_t f = a.fanout, fs = f | 1u;
  const size_t smem = ((size_t)3 * NT + ((NT * fs + 3u) & ~3u) + 2 * (size_t)a.hslots * a.nfy) * sizeof(uint32_t);
  auto kern = fc_sample_kernel<NT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const int occ = occupancy(kern, NT, smem);
  uint64_t cap = (uint64_t)sm_count() * occ / K;
  if (cap < 1) cap = 1;
  uint64_t gx = ((uint64_t)a.n_max + NT - 1) / NT;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  kern<<<dim3((unsigned)gx, K), NT, smem, st>>>(a);
  note_launch();
  return check_last();
}

}  // namespace

bool fast_chain_supported(const fgnn_sample_plan *pl) {
  if (pl->sample_type != 5) return false;
  for (uint32_t i = 0; i < pl->num_layers; ++i) {
    if (pl->fanout[i] == 0 || pl->fanout[i] > 128) return false;
    const uint64_t slots = (uint64_t)pl->in_max[i] * pl->fanout[i];
    if (slots > (uint64_t)kMaxChainCtas * kFcChunk) return false;  // one aggregate slot per 2048-slot chunk
    if (!pl->pos[i]) return false;
  }
  return pl->capacity <= 0x80000000ull;
}

// K mini-batches through all layers on one stream.  plans[k] / outs[k] may differ in table, scratch, outputs and
// version only; topology, fanouts and sampler must be shared (checked by the caller in batch.cu).
int fast_chain_launch(const fgnn_sample_plan *const *plans, const fgnn_sample_out *const *outs,
                      const uint32_t *const *seeds, const uint32_t *n_seeds_max,
                      const uint32_t *const *d_n_seeds, const uint64_t *batch_keys, uint32_t K, cudaStream_t st) {
  const fgnn_sample_plan *p0 = plans[0];
  const uint32_t L = p0->num_layers;
  // versioned tables need every padded index and local id below 2^24 (the tag lives in bits 24..30)
  bool versioned = true;
  for (uint32_t k = 0; k < K; ++k)
    if (plans[k]->version == 0 || plans[k]->version > 126) versioned = false;
  uint64_t max_nodes = p0->in_max[0] + (uint64_t)p0->in_max[0] * p0->fanout[0];
  for (uint32_t i = 0; i < L; ++i)
    if ((uint64_t)p0->in_max[i] * p0->fanout[i] >= (1u << 24)) versioned = false;
  if (max_nodes >= (1u << 24)) versioned = false;
  if (!versioned) {
    for (uint32_t k = 0; k < K; ++k) {
      cudaError_t e = cudaMemsetAsync(plans[k]->table, 0xFF, fgnn_k_ht_bytes(plans[k]->capacity), st);
      if (e != cudaSuccess) return (int)e;
    }
    trace_mark(st, FGNN_TRACE_TABLE_RESET);
  }
  FcArgs a;
  memset(&a, 0, sizeof(a));
  a.indptr = p0->indptr;
  a.indices = p0->indices;
  a.mask = (uint32_t)(p0->capacity - 1);
  a.vmask = versioned ? kFcVerMask : 0u;
  a.valmask = versioned ? 0x00FFFFFFu : 0x7FFFFFFFu;
  for (int i = (int)L - 1; i >= 0; --i) {  // cuda_loops.cc:87
    a.layer = (uint32_t)i;
    a.first = (i == (int)L - 1) ? 1u : 0u;
    a.fanout = p0->fanout[i];
    a.n_max = p0->in_max[i];
    uint32_t n_launch = 0;
    for (uint32_t k = 0; k < K; ++k) {
      FcBatch &b = a.b[k];
      b.seeds = seeds[k];
      b.d_n_seeds = d_n_seeds ? d_n_seeds[k] : nullptr;
      b.n2o = outs[k]->n2o;
      b.table = (Bucket *)plans[k]->table;
      b.num_items = plans[k]->num_items;
      b.counts = outs[k]->counts;
      // ONE padded scratch array per layer: the sampler writes the picked ids into it, the insert replaces every id
      // by its bucket position in place (each slot is read and rewritten by the same thread)
      b.dst = plans[k]->pos[i];
      b.pos = plans[k]->pos[i];
      b.row = outs[k]->row[i];
      b.col = outs[k]->col[i];
      b.ws = (ChainWs *)plans[k]->chain_ws;
      b.key = make_rng_key(fgnn_rng{plans[k]->seed, batch_keys[k], (uint32_t)i});
      b.n_seed_max = n_seeds_max[k];
      b.vtag = versioned ? (plans[k]->version << 24) : 0u;
      if (n_seeds_max[k] > n_launch) n_launch = n_seeds_max[k];
    }
    if (a.first) a.n_max = n_launch;  // the seed count is known on the host: no CTAs for absent seeds
    if (a.n_max == 0) {
      // an empty mini-batch still needs its counts: the kernels below handle n == 0 with one CTA
      a.n_max = 1;
    }
    // Fisher-Yates map: power-of-two slots >= 2 * fanout per lane; as many lanes as fit 16 KB
    uint32_t H = 8;
    while (H < 2 * a.fanout) H <<= 1;
    a.hslots = H;
    a.hshift = 32;
    for (uint32_t h = H; h > 1; h >>= 1) --a.hshift;
    // tile size: small layers are cut into 64-seed tiles so that they still spread over the chip
    const bool small = (uint64_t)a.n_max * K <= (uint64_t)sm_count() * 256ull;
    const uint32_t NT = small ? 64u : 128u;
    uint32_t nfy = NT;
    while (nfy > 32 && (size_t)2 * H * nfy * 4 > 16 * 1024) nfy >>= 1;
    a.nfy = nfy;
    int rc = small ? launch_sample<64>(a, K, st) : launch_sample<128>(a, K, st);
    if (rc) return rc;
    trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_SAMPLE);
    {
      static const int occ_ins = occupancy(fc_insert_kernel, kBlock, 0);
      const uint64_t items = (uint64_t)a.n_max * (a.fanout + (a.first ? 1u : 0u));
      uint64_t gi = (items + (uint64_t)kBlock * kFcInsIlp - 1) / ((uint64_t)kBlock * kFcInsIlp);
      uint64_t cap_i = (uint64_t)sm_count() * occ_ins / K;
      if (cap_i < 1) cap_i = 1;
      if (gi > cap_i) gi = cap_i;
      if (gi < 1) gi = 1;
      fc_insert_kernel<<<dim3((unsigned)gi, K), kBlock, 0, st>>>(a);
      note_launch();
      rc = check_last();
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_INSERT);
    }
    if (a.first) a.n_max = p0->in_max[i];  // the compaction bounds by the plan (counts clamp it)
    const uint64_t slots = (uint64_t)a.n_max * a.fanout;
    uint64_t P = (slots + kFcChunk - 1) / kFcChunk;
    if (a.first) P = ((uint64_t)(n_launch ? n_launch : 1) * a.fanout + kFcChunk - 1) / kFcChunk;
    if (P < 1) P = 1;
    fc_compact_kernel<<<dim3((unsigned)P, K), kBlock, 0, st>>>(a);
    note_launch();
    rc = check_last();
    if (rc) return rc;
    trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_COMPACT);
  }
  return 0;
}

}  // namespace fgnn
