// OrderedHashTable: running set of unique global ids with stable local
// numbering (reference: cuda_hashtable.cu / cuda_hashtable.h, cuda_mapping.cu).
//
// B200-first layout: one 8-byte bucket {key, local} (the reference uses a
// 16-byte {key, local, index, version} bucket), open addressing with linear
// probing over a power-of-two table sized to stay mostly L2-resident.  The
// "which duplicate owns the id" race of the reference (atomicCAS winner writes
// its input index, cuda_hashtable.cu:49-61) is replaced by a deterministic
// rule: while an id is new in this fill its `local` word holds
// 0x80000000|min(input index) maintained with atomicMin; assigned local ids
// (< 2^31) compare smaller, so the same atomicMin leaves ids from earlier fills
// untouched.  Fill remembers each item's bucket position so that remapping the
// edge list afterwards is a direct 4-byte read instead of a second probe.
#include "hashtable.cuh"

namespace fgnn {
namespace {

// returns bucket position holding `id` (inserting it if absent)
__device__ __forceinline__ uint32_t insert_key(Bucket *table, uint32_t mask, uint32_t id) {
  uint32_t pos = hash_id(id, mask);
  while (true) {
    const uint32_t cur = table[pos].key;
    if (cur == id) return pos;
    if (cur == kEmpty) {
      const uint32_t old = atomicCAS(&table[pos].key, kEmpty, id);
      if (old == kEmpty || old == id) return pos;
    }
    pos = (pos + 1) & mask;
  }
}

__global__ void __launch_bounds__(kBlock)
ht_fill_unique_kernel(Bucket *table, uint32_t mask, const uint32_t *__restrict__ input,
                      uint32_t n_max, const uint32_t *__restrict__ d_n, uint32_t *n2o,
                      uint32_t *d_num_items, int is_last_writer) {
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t base = *d_num_items;
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const uint32_t id = __ldg(input + i);
    const uint32_t pos = insert_key(table, mask, id);
    // ids are unique by contract: local = offset + index (cuda_hashtable.cu:164-170)
    table[pos].local = base + i;
    n2o[base + i] = id;
  }
  (void)is_last_writer;
}

__global__ void ht_bump_kernel(uint32_t *d_num_items, uint32_t n_max, const uint32_t *d_n) {
  *d_num_items += load_count(n_max, d_n);
}

constexpr int kInsIlp = 4;

// ncu r1_a: 48 us for 0.85 M items with one dependent probe per thread and an
// unconditional atomicMin (hub ids serialise in the L2 atomic unit).  Now: four
// independent first probes in flight per thread, and the atomicMin is skipped
// whenever the snapshot already shows a smaller owner index / an assigned id.
__global__ void __launch_bounds__(kBlock)
ht_insert_kernel(Bucket *table, uint32_t mask, const uint32_t *__restrict__ input,
                 uint32_t n_max, const uint32_t *__restrict__ d_n, uint32_t mult,
                 uint32_t *__restrict__ pos_out) {
  // mult > 1: `input` is a padded [seed][mult] block (EMPTY = hole) and *d_n counts seeds
  const uint32_t n = mult > 1 ? load_count(n_max / mult, d_n) * mult : load_count(n_max, d_n);
  const uint32_t stride = gridDim.x * kBlock;
  for (uint32_t i0 = blockIdx.x * kBlock + threadIdx.x; i0 < n; i0 += stride * kInsIlp) {
    uint32_t id[kInsIlp], pos[kInsIlp];
    uint2 b[kInsIlp];
#pragma unroll
    for (int u = 0; u < kInsIlp; ++u) {
      const uint32_t i = i0 + u * stride;
      id[u] = i < n ? __ldg(input + i) : kEmpty;
      pos[u] = hash_id(id[u], mask);
    }
#pragma unroll
    for (int u = 0; u < kInsIlp; ++u)
      if (id[u] != kEmpty) b[u] = load_bucket(table + pos[u]);
#pragma unroll
    for (int u = 0; u < kInsIlp; ++u) {
      const uint32_t i = i0 + u * stride;
      if (id[u] != kEmpty) {  // in range and not a hole
        pos_out[i] = insert_item(table, mask, id[u], i, pos[u], b[u]);
      }
    }
  }
}

struct CompactSmem {
  uint32_t warp[kBlock / 32 + 1];
  ChainSmem chain;
};

// wait until the owner of bucket `bp` has received its local id (only lower
// tickets or this CTA's earlier tiles can own it, see ht_compact_kernel)
__device__ __forceinline__ uint32_t wait_local(const Bucket *table, uint32_t bp) {
  uint32_t w = ld_relaxed_u32(&table[bp].local);
  while (w & kPending) {
    __nanosleep(32);
    w = ld_relaxed_u32(&table[bp].local);
  }
  return w;
}

// Assign local ids to the ids that are new in this fill (first occurrence order)
// and — when out_local != nullptr — remap EVERY item of the fill to its local id
// in the same launch (the reference runs FillWithDuplicates, then GPUMapEdges as
// a second probe pass, cuda_hashtable.cu:725-807 + cuda_mapping.cu:68-81).
//
// Remap without a grid-wide barrier: the owner of an id is its smallest input
// index, so it always lives in this CTA's chunk at an earlier position or in a
// chunk with a LOWER ticket.  Lower tickets never wait on higher ones (they only
// look back), so spinning on the bucket's local word until the pending bit
// clears cannot deadlock; the CTA assigns all of its own owners first and only
// then waits, so it never delays the tickets that wait on it.
__global__ void __launch_bounds__(kBlock)
ht_compact_kernel(Bucket *table, const uint32_t *__restrict__ input, uint32_t n_max,
                  const uint32_t *__restrict__ d_n, const uint32_t *__restrict__ pos,
                  uint32_t *__restrict__ n2o, uint32_t *d_num_items,
                  uint32_t *__restrict__ out_local, uint32_t *count_copy, uint32_t *count_copy2,
                  ChainWs *ws) {
  __shared__ CompactSmem sm;
  const uint32_t n = load_count(n_max, d_n);
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, kBlock, &begin, &end);
  const uint32_t items0 = *d_num_items;  // stable: only the last finisher updates it, at the end

  // bucket + local word of the first kCache tiles of the chunk stay in registers so
  // the random bucket read happens once (a chunk is <= 4 tiles up to ~1.2 M items)
  constexpr int kCache = 4;
  uint32_t bpr[kCache], wr[kCache];
  unsigned long long partial = 0;
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const uint32_t i = begin + it * kBlock + threadIdx.x;
    bpr[it] = 0;
    wr[it] = 0;
    if (i < end) {
      bpr[it] = pos[i];
      wr[it] = table[bpr[it]].local;
      partial += (wr[it] == (kPending | i)) ? 1u : 0u;
    }
  }
  for (uint32_t i = begin + kCache * kBlock + threadIdx.x; i < end; i += kBlock)
    partial += (table[pos[i]].local == (kPending | i)) ? 1u : 0u;
  unsigned long long chunk_total;
  unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);

#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const uint32_t t0 = begin + it * kBlock;
    if (t0 < end) {  // uniform across the CTA
      const uint32_t i = t0 + threadIdx.x;
      const uint32_t flag = (i < end && wr[it] == (kPending | i)) ? 1u : 0u;
      uint32_t tile_total;
      const uint32_t excl = block_excl_scan(flag, sm.warp, &tile_total);
      if (flag) {
        const uint32_t local = items0 + (uint32_t)base + excl;
        table[bpr[it]].local = local;
        n2o[local] = __ldg(input + i);
        wr[it] = local;
      }
      base += tile_total;
    }
  }
  for (uint32_t t0 = begin + kCache * kBlock; t0 < end; t0 += kBlock) {
    const uint32_t i = t0 + threadIdx.x;
    uint32_t flag = 0, bp = 0, w = 0;
    if (i < end) {
      bp = pos[i];
      w = table[bp].local;
      flag = (w == (kPending | i)) ? 1u : 0u;
    }
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(flag, sm.warp, &tile_total);
    if (flag) {
      w = items0 + (uint32_t)base + excl;
      table[bp].local = w;
      n2o[w] = __ldg(input + i);
    }
    base += tile_total;
  }
  if (out_local) {
    __syncthreads();  // this CTA's owners are all assigned and visible
#pragma unroll
    for (int it = 0; it < kCache; ++it) {
      const uint32_t i = begin + it * kBlock + threadIdx.x;
      if (i < end) out_local[i] = (wr[it] & kPending) ? wait_local(table, bpr[it]) : wr[it];
    }
    // long chunks: second walk, only now may the CTA wait on lower tickets (waiting inside the loop above
    // delayed this CTA's own owners and, through them, every higher ticket: ncu r1_q, 24 -> 156 us)
    for (uint32_t i = begin + kCache * kBlock + threadIdx.x; i < end; i += kBlock)
      out_local[i] = wait_local(table, pos[i]);
  }
  // the item counter is only advanced by the CTA that finishes LAST (every other CTA has read items0 by
  // then); the number of new ids = base at the end of the last chunk travels through the pad word
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p == gridDim.x - 1) ws->pad[0] = (uint32_t)base;
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    sm.chain.last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sm.chain.last) {
    for (uint32_t t = threadIdx.x; t < gridDim.x; t += kBlock) ws->agg[t] = 0ull;
    if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t total = items0 + *((volatile unsigned int *)&ws->pad[0]);
      *d_num_items = total;
      if (count_copy) *count_copy = total;
      if (count_copy2) *count_copy2 = total;
      ws->pad[0] = 0u;
      ws->ticket = 0u;
      ws->done = 0u;
    }
  }
}

// FillWithDuplicates (second half) + compact_edge + GPUMapEdges for a PADDED sampler output
// (sample_khop2_pad_kernel): item i = seed i / fanout, pick i % fanout, EMPTY = hole.  ONE chained scan carries
// both running counts, packed {valid edges : 31 | new ids : 31}; per valid item the kernel writes the compact COO
//   col[e] = i / fanout            (the seed's local id: seeds are the first entries of the unique list)
//   row[e] = local id of dst[i]    (owners first, then everyone reads its bucket; see ht_compact_kernel)
// in seed-major order, exactly what sample + count_edge + compact_edge + FillWithDuplicates + MapEdges of the
// reference produce (cuda_sampling_khop2.cu:121-175, cuda_hashtable.cu:725-807, cuda_mapping.cu:68-81).
__global__ void __launch_bounds__(kBlock, 4)
ht_compact_pad_kernel(Bucket *table, const uint32_t *__restrict__ dst, uint32_t n_seed_max,
                      const uint32_t *__restrict__ d_n_seed, uint32_t fanout,
                      const uint32_t *__restrict__ pos, uint32_t *__restrict__ n2o, uint32_t *d_num_items,
                      uint32_t *__restrict__ out_row, uint32_t *__restrict__ out_col, uint32_t *count_edge,
                      uint32_t *count_src, uint32_t *count_next, ChainWs *ws) {
  __shared__ CompactSmem sm;
  const uint32_t n = load_count(n_seed_max, d_n_seed) * fanout;
  const uint32_t p = chain_ticket(ws, &sm.chain);
  uint32_t begin, end;
  chunk_range(n, p, gridDim.x, kBlock, &begin, &end);
  const uint32_t items0 = *d_num_items;  // stable: only the last finisher updates it, at the end

  constexpr int kCache = 6;  // one resident wave (>= 4 CTAs/SM = 592 tickets) covers ~0.9 M items in <= 6 tiles
  uint32_t bpr[kCache], wr[kCache], idr[kCache];
  unsigned long long partial = 0;
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const uint32_t i = begin + it * kBlock + threadIdx.x;
    bpr[it] = 0;
    wr[it] = 0;
    idr[it] = kEmpty;
    if (i < end) idr[it] = __ldg(dst + i);
    if (idr[it] != kEmpty) bpr[it] = pos[i];
  }
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const uint32_t i = begin + it * kBlock + threadIdx.x;
    if (idr[it] != kEmpty) {
      wr[it] = table[bpr[it]].local;
      partial += (1ull << 31) + ((wr[it] == (kPending | i)) ? 1ull : 0ull);
    }
  }
  for (uint32_t i = begin + kCache * kBlock + threadIdx.x; i < end; i += kBlock) {
    if (__ldg(dst + i) != kEmpty)
      partial += (1ull << 31) + ((table[pos[i]].local == (kPending | i)) ? 1ull : 0ull);
  }
  unsigned long long chunk_total;
  const unsigned long long base = chain_scan(ws, &sm.chain, p, partial, &chunk_total);
  uint32_t base_edge = (uint32_t)(base >> 31), base_new = (uint32_t)(base & 0x7FFFFFFFull);

  uint32_t eoff[kCache];
#pragma unroll
  for (int it = 0; it < kCache; ++it) {
    const uint32_t t0 = begin + it * kBlock;
    eoff[it] = 0;
    if (t0 < end) {  // uniform across the CTA
      const uint32_t i = t0 + threadIdx.x;
      const uint32_t valid = idr[it] != kEmpty ? 1u : 0u;
      const uint32_t isnew = (valid && wr[it] == (kPending | i)) ? 1u : 0u;
      uint32_t tile_total;
      const uint32_t excl = block_excl_scan((valid << 16) | isnew, sm.warp, &tile_total);
      if (isnew) {
        const uint32_t local = items0 + base_new + (excl & 0xFFFFu);
        table[bpr[it]].local = local;
        n2o[local] = idr[it];
        wr[it] = local;
      }
      eoff[it] = base_edge + (excl >> 16);
      if (valid) out_col[eoff[it]] = i / fanout;
      base_edge += tile_total >> 16;
      base_new += tile_total & 0xFFFFu;
    }
  }
  const uint32_t edge_after_cache = base_edge;
  for (uint32_t t0 = begin + kCache * kBlock; t0 < end; t0 += kBlock) {  // long chunks, first walk: owners + col
    const uint32_t i = t0 + threadIdx.x;
    uint32_t valid = 0, isnew = 0, bp = 0, id = kEmpty;
    if (i < end) id = __ldg(dst + i);
    if (id != kEmpty) {
      valid = 1;
      bp = pos[i];
      isnew = (table[bp].local == (kPending | i)) ? 1u : 0u;
    }
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan((valid << 16) | isnew, sm.warp, &tile_total);
    if (isnew) {
      const uint32_t local = items0 + base_new + (excl & 0xFFFFu);
      table[bp].local = local;
      n2o[local] = id;
    }
    if (valid) out_col[base_edge + (excl >> 16)] = i / fanout;
    base_edge += tile_total >> 16;
    base_new += tile_total & 0xFFFFu;
  }
  __syncthreads();  // this CTA's owners are all assigned and visible
#pragma unroll
  for (int it = 0; it < kCache; ++it)
    if (idr[it] != kEmpty) out_row[eoff[it]] = (wr[it] & kPending) ? wait_local(table, bpr[it]) : wr[it];
  // long chunks, second walk: only now may the CTA wait on lower tickets
  uint32_t e_walk = edge_after_cache;
  for (uint32_t t0 = begin + kCache * kBlock; t0 < end; t0 += kBlock) {
    const uint32_t i = t0 + threadIdx.x;
    const uint32_t valid = (i < end && __ldg(dst + i) != kEmpty) ? 1u : 0u;
    uint32_t tile_total;
    const uint32_t excl = block_excl_scan(valid, sm.warp, &tile_total);
    if (valid) out_row[e_walk + excl] = wait_local(table, pos[i]);
    e_walk += tile_total;
  }

  // totals travel through the pad words; the CTA that finishes LAST publishes them (every other CTA has
  // read items0 by then) and re-arms the workspace
  __syncthreads();
  if (threadIdx.x == 0) {
    if (p == gridDim.x - 1) {
      ws->pad[0] = base_new;
      ws->pad[1] = base_edge;
    }
    __threadfence();
    const unsigned int prev = atomicAdd(&ws->done, 1u);
    sm.chain.last = (prev == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sm.chain.last) {
    for (uint32_t t = threadIdx.x; t < gridDim.x; t += kBlock) ws->agg[t] = 0ull;
    if (threadIdx.x == 0) {
      __threadfence();
      const uint32_t total = items0 + *((volatile unsigned int *)&ws->pad[0]);
      const uint32_t edges = *((volatile unsigned int *)&ws->pad[1]);
      *d_num_items = total;
      if (count_edge) *count_edge = edges;
      if (count_src) *count_src = total;
      if (count_next) *count_next = total;
      ws->pad[0] = 0u;
      ws->pad[1] = 0u;
      ws->ticket = 0u;
      ws->done = 0u;
    }
  }
}

// FillWithUnique into an empty table (first fill of a batch): local id = index, count = n
__global__ void __launch_bounds__(kBlock)
ht_fill_unique_first_kernel(Bucket *table, uint32_t mask, const uint32_t *__restrict__ input,
                            uint32_t n_max, const uint32_t *__restrict__ d_n, uint32_t *n2o,
                            uint32_t *d_num_items, uint32_t *count_copy) {
  const uint32_t n = load_count(n_max, d_n);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *d_num_items = n;
    if (count_copy) *count_copy = n;
  }
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    const uint32_t id = __ldg(input + i);
    const uint32_t pos = insert_key(table, mask, id);
    table[pos].local = i;
    n2o[i] = id;
  }
}

__global__ void __launch_bounds__(kBlock)
ht_map_kernel(const Bucket *__restrict__ table, uint32_t mask, const uint32_t *__restrict__ global,
              const uint32_t *__restrict__ pos, uint32_t n_max, const uint32_t *__restrict__ d_n,
              uint32_t *__restrict__ out_local) {
  const uint32_t n = load_count(n_max, d_n);
  for (uint32_t i = blockIdx.x * kBlock + threadIdx.x; i < n; i += gridDim.x * kBlock) {
    uint32_t local;
    if (pos) {
      local = table[pos[i]].local;
    } else {
      const uint32_t id = __ldg(global + i);
      uint32_t b = hash_id(id, mask);
      local = kEmpty;
      while (true) {  // SearchForPositionO2N, cuda_hashtable.h:69-83
        const Bucket bk = table[b];
        if (bk.key == id) { local = bk.local; break; }
        if (bk.key == kEmpty) break;  // absent -> EMPTY (the reference asserts)
        b = (b + 1) & mask;
      }
    }
    out_local[i] = local;
  }
}

}  // namespace

int ht_insert_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                     const uint32_t *d_n, uint32_t *pos, cudaStream_t st, uint32_t mult) {
  static const int occ1 = occupancy(ht_insert_kernel, kBlock, 0);
  const int grid1 = persistent_grid(n_max, kBlock * kInsIlp, occ1, false);
  ht_insert_kernel<<<grid1, kBlock, 0, st>>>((Bucket *)table, (uint32_t)(capacity - 1), input, n_max, d_n,
                                            mult ? mult : 1u, pos);
  note_launch();
  return check_last();
}

int ht_compact_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                      const uint32_t *d_n, const uint32_t *pos, uint32_t *n2o, uint32_t *d_num_items,
                      uint32_t *out_local, uint32_t *count_copy, uint32_t *count_copy2, void *chain_ws,
                      cudaStream_t st) {
  (void)capacity;
  static const int occ2 = occupancy(ht_compact_kernel, kBlock, 0);
  const int grid2 = persistent_grid(n_max, kBlock, occ2, true);
  ht_compact_kernel<<<grid2, kBlock, 0, st>>>((Bucket *)table, input, n_max, d_n, pos, n2o, d_num_items,
                                             out_local, count_copy, count_copy2, (ChainWs *)chain_ws);
  note_launch();
  return check_last();
}

int ht_compact_pad_launch(void *table, const uint32_t *dst, uint32_t n_seed_max, const uint32_t *d_n_seed,
                          uint32_t fanout, const uint32_t *pos, uint32_t *n2o, uint32_t *d_num_items,
                          uint32_t *out_row, uint32_t *out_col, uint32_t *count_edge, uint32_t *count_src,
                          uint32_t *count_next, void *chain_ws, cudaStream_t st) {
  // ONE resident wave: tickets beyond residency start a second round of the whole latency chain (measured r1_q c6:
  // 1024 tickets on 740 resident slots, 56 us instead of 24 us)
  static const int occ = occupancy(ht_compact_pad_kernel, kBlock, 0);
  const int grid = persistent_grid((uint64_t)n_seed_max * fanout, kBlock, occ, true);
  ht_compact_pad_kernel<<<grid, kBlock, 0, st>>>((Bucket *)table, dst, n_seed_max, d_n_seed, fanout, pos, n2o,
                                                 d_num_items, out_row, out_col, count_edge, count_src,
                                                 count_next, (ChainWs *)chain_ws);
  note_launch();
  return check_last();
}

int ht_fill_unique_first_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                                const uint32_t *d_n, uint32_t *n2o, uint32_t *d_num_items,
                                uint32_t *count_copy, cudaStream_t st) {
  const int grid = persistent_grid(n_max ? n_max : 1, kBlock, 8, false);
  ht_fill_unique_first_kernel<<<grid, kBlock, 0, st>>>((Bucket *)table, (uint32_t)(capacity - 1), input,
                                                      n_max, d_n, n2o, d_num_items, count_copy);
  note_launch();
  return check_last();
}

}  // namespace fgnn

using namespace fgnn;

extern "C" size_t fgnn_k_ht_capacity(size_t max_items) {
  // power of two >= 1.5 * max_items, at least 1024 buckets
  size_t want = max_items + (max_items >> 1) + 1;
  size_t cap = 1024;
  while (cap < want) cap <<= 1;
  return cap;
}

extern "C" size_t fgnn_k_ht_bytes(size_t capacity) { return capacity * sizeof(Bucket); }

extern "C" int fgnn_k_ht_reset(void *table, size_t capacity, uint32_t *d_num_items,
                               fgnn_stream_t stream) {
  if (!table || !d_num_items || (capacity & (capacity - 1))) return FGNN_ERR_BAD_ARG;
  cudaError_t e = cudaMemsetAsync(table, 0xFF, capacity * sizeof(Bucket), (cudaStream_t)stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(d_num_items, 0, sizeof(uint32_t), (cudaStream_t)stream);
  return (int)e;
}

extern "C" int fgnn_k_ht_fill_unique(void *table, size_t capacity, const uint32_t *input,
                                     uint32_t n_max, const uint32_t *d_n, uint32_t *n2o,
                                     uint32_t *d_num_items, fgnn_stream_t stream) {
  if (!table || !n2o || !d_num_items || (capacity & (capacity - 1))) return FGNN_ERR_BAD_ARG;
  if (capacity > 0x80000000ull) return FGNN_ERR_UNSUPPORTED;
  if (n_max == 0) return 0;
  if (!input) return FGNN_ERR_BAD_ARG;
  const int grid = persistent_grid(n_max, kBlock, 8, false);
  ht_fill_unique_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>(
      (Bucket *)table, (uint32_t)(capacity - 1), input, n_max, d_n, n2o, d_num_items, 0);
  ht_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_num_items, n_max, d_n);
  note_launch(2);
  return check_last();
}

static int check_fill_args(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                           uint32_t *pos, uint32_t *n2o, uint32_t *d_num_items, void *chain_ws) {
  if (!table || !n2o || !d_num_items || !chain_ws || (capacity & (capacity - 1))) return FGNN_ERR_BAD_ARG;
  if (capacity > 0x80000000ull || n_max >= kPending) return FGNN_ERR_UNSUPPORTED;
  if (n_max > 0 && (!input || !pos)) return FGNN_ERR_BAD_ARG;
  return 0;
}

extern "C" int fgnn_k_ht_fill_duplicates(void *table, size_t capacity, const uint32_t *input,
                                         uint32_t n_max, const uint32_t *d_n, uint32_t *pos,
                                         uint32_t *n2o, uint32_t *d_num_items, void *chain_ws,
                                         fgnn_stream_t stream) {
  if (int rc = check_fill_args(table, capacity, input, n_max, pos, n2o, d_num_items, chain_ws)) return rc;
  if (n_max == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ht_insert_launch(table, capacity, input, n_max, d_n, pos, st, 1)) return rc;
  return ht_compact_launch(table, capacity, input, n_max, d_n, pos, n2o, d_num_items, nullptr, nullptr,
                           nullptr, chain_ws, st);
}

extern "C" int fgnn_k_ht_fill_duplicates_map(void *table, size_t capacity, const uint32_t *input,
                                             uint32_t n_max, const uint32_t *d_n, uint32_t *pos,
                                             uint32_t *n2o, uint32_t *d_num_items, uint32_t *out_local,
                                             void *chain_ws, fgnn_stream_t stream) {
  if (int rc = check_fill_args(table, capacity, input, n_max, pos, n2o, d_num_items, chain_ws)) return rc;
  if (n_max == 0) return 0;
  if (!out_local) return FGNN_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = ht_insert_launch(table, capacity, input, n_max, d_n, pos, st, 1)) return rc;
  return ht_compact_launch(table, capacity, input, n_max, d_n, pos, n2o, d_num_items, out_local, nullptr,
                           nullptr, chain_ws, st);
}

extern "C" int fgnn_k_ht_map(const void *table, size_t capacity, const uint32_t *global,
                             const uint32_t *pos, uint32_t n_max, const uint32_t *d_n,
                             uint32_t *out_local, fgnn_stream_t stream) {
  if (!table || !out_local || (capacity & (capacity - 1))) return FGNN_ERR_BAD_ARG;
  if (n_max == 0) return 0;
  if (!global && !pos) return FGNN_ERR_BAD_ARG;
  static const int occ = occupancy(ht_map_kernel, kBlock, 0);
  const int grid = persistent_grid(n_max, kBlock, occ, false);
  ht_map_kernel<<<grid, kBlock, 0, (cudaStream_t)stream>>>((const Bucket *)table,
                                                          (uint32_t)(capacity - 1), global, pos,
                                                          n_max, d_n, out_local);
  note_launch();
  return check_last();
}
