// OrderedHashTable device primitives shared by hashtable.cu (stand-alone fill /
// map kernels) and the samplers that insert their picks while they gather them.
// Layout and ownership rule: see hashtable.cu.
#pragma once
#include "common.cuh"

namespace fgnn {

struct __align__(8) Bucket {
  uint32_t key;
  uint32_t local;
};

constexpr uint32_t kPending = 0x80000000u;

__device__ __forceinline__ uint32_t hash_id(uint32_t id, uint32_t mask) {
  return ((id * 0x9E3779B1u) >> 7) & mask;
}

__device__ __forceinline__ uint2 load_bucket(const Bucket *b) {
  return *reinterpret_cast<const uint2 *>(b);
}

// finish the probe sequence of `id` starting from a first-probe snapshot `b` of
// bucket `pos`; returns the bucket position, *local_seen = last observed local
__device__ __forceinline__ uint32_t resolve_insert(Bucket *table, uint32_t mask, uint32_t id,
                                                   uint32_t pos, uint2 b, uint32_t *local_seen) {
  while (true) {
    if (b.x == id) { *local_seen = b.y; return pos; }
    if (b.x == kEmpty) {
      const uint32_t old = atomicCAS(&table[pos].key, kEmpty, id);
      if (old == kEmpty || old == id) { *local_seen = kEmpty; return pos; }
    }
    pos = (pos + 1) & mask;
    b = load_bucket(table + pos);
  }
}

// insert item `index` (its position in the fill's input order) holding `id`;
// returns its bucket.  The smallest index of a new id becomes its owner.
__device__ __forceinline__ uint32_t insert_item(Bucket *table, uint32_t mask, uint32_t id,
                                                uint32_t index, uint32_t pos, uint2 first) {
  uint32_t seen;
  const uint32_t bp = resolve_insert(table, mask, id, pos, first, &seen);
  if (seen > (kPending | index)) atomicMin(&table[bp].local, kPending | index);
  return bp;
}

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// spin until the owner of bucket `bp` has received its local id; returns the whole local word
__device__ __forceinline__ uint32_t wait_local_word(const Bucket *table, uint32_t bp) {
  uint32_t w = ld_relaxed_u32(&table[bp].local);
  while (w & kPending) {
    __nanosleep(32);
    w = ld_relaxed_u32(&table[bp].local);
  }
  return w;
}

// optional "insert while sampling" target handed to the sampler kernels
struct HtInsert {
  Bucket *table;      // nullptr: do not insert
  uint32_t mask;
  uint32_t *pos_out;  // bucket of every emitted item, in output order
};

// internal launchers used by batch.cu (fgnn_k_sample_batch)
int sample_khop_launch(int variant, const uint32_t *indptr, const uint32_t *indices,
                       const uint32_t *input, uint32_t n_max, const uint32_t *d_n, uint32_t fanout,
                       fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst, uint32_t *out_src_local,
                       uint32_t *d_num_out, void *chain_ws, HtInsert ht, cudaStream_t stream);
// compact (+ optional remap of the items to local ids, + optional copies of the new item count)
int ht_compact_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                      const uint32_t *d_n, const uint32_t *pos, uint32_t *n2o, uint32_t *d_num_items,
                      uint32_t *out_local, uint32_t *count_copy, uint32_t *count_copy2, void *chain_ws,
                      cudaStream_t stream);
// mult > 1: `input` is a padded [seed][mult] block with EMPTY holes and *d_n counts seeds
int ht_insert_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                     const uint32_t *d_n, uint32_t *pos, cudaStream_t stream, uint32_t mult = 1);
// padded k-hop sampling (no compaction) + the compaction pass that also assigns the new ids and remaps
int sample_khop2_pad_launch(const uint32_t *indptr, const uint32_t *indices, const uint32_t *input,
                            uint32_t n_max, const uint32_t *d_n, uint32_t fanout, fgnn_rng rng,
                            uint32_t *out_dst_padded, cudaStream_t stream);
int ht_compact_pad_launch(void *table, const uint32_t *dst, uint32_t n_seed_max, const uint32_t *d_n_seed,
                          uint32_t fanout, const uint32_t *pos, uint32_t *n2o, uint32_t *d_num_items,
                          uint32_t *out_row, uint32_t *out_col, uint32_t *count_edge, uint32_t *count_src,
                          uint32_t *count_next, void *chain_ws, cudaStream_t stream);
// FillWithUnique into an EMPTY table: base is 0 by construction, so the count needs no second kernel
int ht_fill_unique_first_launch(void *table, size_t capacity, const uint32_t *input, uint32_t n_max,
                                const uint32_t *d_n, uint32_t *n2o, uint32_t *d_num_items,
                                uint32_t *count_copy, cudaStream_t stream);

// the two-launch-per-layer uniform k-hop chain for up to FGNN_MAX_SUPER mini-batches (fast_chain.cu)
bool fast_chain_supported(const fgnn_sample_plan *pl);
int fast_chain_launch(const fgnn_sample_plan *const *plans, const fgnn_sample_out *const *outs,
                      const uint32_t *const *seeds, const uint32_t *n_seeds_max,
                      const uint32_t *const *d_n_seeds, const uint64_t *batch_keys, uint32_t K,
                      cudaStream_t stream);

}  // namespace fgnn
