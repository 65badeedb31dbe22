/* fgnn_kernels.h — thin C-ABI over the hand-written sm_100a kernels of the
 * factored sampling-and-extraction hot path.
 *
 * This is the boundary the C++ host runtime (csrc/runtime) and the tests call
 * through: plain pointers and sizes, no C++/torch types, no allocation inside,
 * every call asynchronous on `stream`.  Each entry point cites the reference
 * function it replaces (paths relative to /root/reference/samgraph/common/).
 *
 * Conventions
 *   - ids are uint32_t (IdType, common.h:35); FGNN_EMPTY == Constant::kEmptyKey.
 *   - "device counts": an element count is passed as (n_max, d_n).  If d_n is
 *     NULL the count is n_max; otherwise the kernel reads *d_n (<= n_max) on the
 *     device, so a whole mini-batch can be enqueued (or graph-captured) without
 *     a host round trip per layer (the reference syncs ~12x per layer,
 *     cuda_loops.cc:163-166, cuda_hashtable.cu:783-785).
 *   - `chain_ws` is FGNN_CHAIN_WS_BYTES of device memory, zeroed once at
 *     allocation.  Kernels that compact leave it zeroed, so one buffer per
 *     stream can be reused by every call on that stream.
 *   - return value: 0 on success, otherwise a cudaError_t (>0) or a negative
 *     FGNN_ERR_* code.  Nothing falls back to the CPU.
 */
#ifndef FGNN_KERNELS_H
#define FGNN_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGNN_EMPTY 0xFFFFFFFFu
#define FGNN_CHAIN_WS_BYTES (16u + 8u * 4096u)

#define FGNN_ERR_BAD_ARG (-1)
#define FGNN_ERR_UNSUPPORTED (-2)

typedef struct CUstream_st *fgnn_stream_t; /* == cudaStream_t */

/* RNG stream identity: Philox4x32-10, key=(seed.lo, seed.hi^batch_key.hi),
 * ctr=(draw>>2, item, tag, batch_key.lo).  Replaces GPURandomStates
 * (cuda_random_states.cu:36-108: per-thread XORWOW states, wall-clock seed). */
typedef struct fgnn_rng {
  uint64_t seed;
  uint64_t batch_key; /* Engine::GetBatchKey: epoch*steps+step, engine.h:49-53 */
  uint32_t tag;       /* layer index */
} fgnn_rng;

const char *fgnn_k_version(void);
const char *fgnn_k_error_string(int code);
/* number of kernel launches issued through this library since load */
uint64_t fgnn_k_launch_count(void);

/* ---- epoch shuffle --------------------------------------------------------- */
/* GPUShuffler::ReShuffle / DistShuffler::ReShuffle (cuda_shuffler.cc:75-126,
 * dist_shuffler.cc:98-151): a fresh permutation of the train set per epoch, a pure
 * function of (seed, epoch) like the reference's dist shuffler (seed = epoch,
 * dist_shuffler.cc:114).  out[rank of key_i] = train_set[i] with
 * key_i = Philox(seed, batch_key=epoch, tag=FGNN_SHUFFLE_TAG, item=i, draw 0),
 * stable ascending sort. */
#define FGNN_SHUFFLE_TAG 0xFFFF0001u
size_t fgnn_k_shuffle_workspace_bytes(size_t n);
int fgnn_k_shuffle(const uint32_t *train_set, size_t n, uint64_t seed,
                   uint64_t epoch, uint32_t *out, void *workspace,
                   size_t workspace_bytes, fgnn_stream_t stream);

/* ---- uniform k-hop sampling --------------------------------------------- */
/* variant 0: GPUSampleKHop0 (cuda_sampling_khop0.cu:178-253, reservoir)
 * variant 2: GPUSampleKHop2 (cuda_sampling_khop2.cu:177-252, Fisher-Yates,
 *            stateless: the CSR is never mutated)
 * Output is the compact COO in seed-major order (== sample + count_edge +
 * compact_edge of the reference, fused):  out_src[e] global seed id (may be
 * NULL), out_dst[e] global neighbour id, out_src_local[e] = index of the seed
 * in `input` (may be NULL; equals the seed's local id, cuda_loops.cc:203-221).
 * *d_num_out <- number of edges.  Outputs must hold n_max*fanout entries. */
int fgnn_k_sample_khop(int variant, const uint32_t *indptr,
                       const uint32_t *indices, const uint32_t *input,
                       uint32_t n_max, const uint32_t *d_n, uint32_t fanout,
                       fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst,
                       uint32_t *out_src_local, uint32_t *d_num_out,
                       void *chain_ws, fgnn_stream_t stream);

/* ---- with-replacement samplers (sort by src id + adjacent dedup) ---------- */
/* kind 1: GPUSampleKHop1            (cuda_sampling_khop1.cu:130-234)
 * kind 2: GPUSampleWeightedKHop     (cuda_sampling_weighted_khop.cu:132-236)
 * kind 4: GPUSampleWeightedKHopPrefix (cuda_sampling_weighted_khop_prefix.cu:137-255)
 * (kind values follow SampleType, common.h:50-58.)  prob/alias/prefix are the
 * per-edge tables of Dataset (common.h:141-144); unused ones may be NULL.
 * Output order follows the reference: rows ordered by ascending seed id, an
 * entry is dropped when equal to its successor.  workspace: see
 * fgnn_k_sample_replace_workspace_bytes. */
size_t fgnn_k_sample_replace_workspace_bytes(uint32_t n_max, uint32_t fanout);
int fgnn_k_sample_replace(int kind, const uint32_t *indptr,
                          const uint32_t *indices, const float *prob_table,
                          const uint32_t *alias_table,
                          const float *prob_prefix_table, const uint32_t *input,
                          uint32_t n_max, const uint32_t *d_n, uint32_t fanout,
                          fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst,
                          uint32_t *out_src_local, uint32_t *d_num_out,
                          void *workspace, size_t workspace_bytes,
                          void *chain_ws, fgnn_stream_t stream);
/* The same sampler with the seeds ordered by a rank-by-bitmap instead of a radix sort (the seeds of a layer are
 * unique: rank(v) = number of seeds below v = popcount of a V-bit seed bitmap below bit v).  `rank_ws` holds
 * fgnn_k_seed_rank_workspace_bytes(num_nodes) bytes that must be ZERO before the first call; every call leaves
 * them zero again.  rank_ws == NULL, num_nodes == 0 or num_nodes > 2^29 select the sort. */
size_t fgnn_k_seed_rank_workspace_bytes(size_t num_nodes);
int fgnn_k_sample_replace_ranked(int kind, const uint32_t *indptr,
                                 const uint32_t *indices, const float *prob_table,
                                 const uint32_t *alias_table,
                                 const float *prob_prefix_table, const uint32_t *input,
                                 uint32_t n_max, const uint32_t *d_n, uint32_t fanout,
                                 fgnn_rng rng, uint32_t *out_src, uint32_t *out_dst,
                                 uint32_t *out_src_local, uint32_t *d_num_out,
                                 void *workspace, size_t workspace_bytes,
                                 void *chain_ws, void *rank_ws, size_t num_nodes,
                                 fgnn_stream_t stream);

/* GPUSampleWeightedKHopHashDedup (cuda_sampling_weighted_khop_hash_dedup.cu:
 * 203-279): alias sampling with rejection until `fanout` distinct; seed-major
 * compact output like fgnn_k_sample_khop.  The reference spins forever when a
 * row has fewer than `fanout` distinct reachable ids; here draws beyond
 * FGNN_HASH_DEDUP_MAX_DRAWS are accepted even if duplicate. */
#define FGNN_HASH_DEDUP_MAX_DRAWS 4096u
int fgnn_k_sample_weighted_hash_dedup(
    const uint32_t *indptr, const uint32_t *indices, const float *prob_table,
    const uint32_t *alias_table, const uint32_t *input, uint32_t n_max,
    const uint32_t *d_n, uint32_t fanout, fgnn_rng rng, uint32_t *out_src,
    uint32_t *out_dst, uint32_t *out_src_local, uint32_t *d_num_out,
    void *chain_ws, fgnn_stream_t stream);

/* ---- PinSAGE random walk + top-K ------------------------------------------ */
/* GPUSampleRandomWalk (cuda_sampling_random_walk.cu:113-161) followed by
 * FrequencyHashmap::GetTopK (cuda_frequency_hashmap.cu:1143-1367), fused: the
 * visit list of a start node never leaves the SM.  out_data[e] = visit count.
 * If tmp_src/tmp_dst are non-NULL the raw walk COO (n_max*W*L entries, layout
 * of random_walk.cu:60-78) is also written, for parity tests.
 * Requires num_walk*walk_len <= 32.  workspace: see *_workspace_bytes. */
size_t fgnn_k_sample_random_walk_workspace_bytes(uint32_t n_max, uint32_t K);
int fgnn_k_sample_random_walk(const uint32_t *indptr, const uint32_t *indices,
                              const uint32_t *input, uint32_t n_max,
                              const uint32_t *d_n, uint32_t walk_len,
                              double restart_prob, uint32_t num_walk,
                              uint32_t K, fgnn_rng rng, uint32_t *out_src,
                              uint32_t *out_dst, uint32_t *out_src_local,
                              uint32_t *out_data, uint32_t *d_num_out,
                              uint32_t *tmp_src, uint32_t *tmp_dst,
                              void *workspace, size_t workspace_bytes,
                              void *chain_ws, fgnn_stream_t stream);

/* ---- OrderedHashTable (cuda_hashtable.cu) ---------------------------------- */
/* Table = `capacity` (power of two) 8-byte buckets {key, local}.  `d_num_items`
 * is the running number of unique ids (device).  `n2o` (>= max items) is the
 * local->global list, i.e. the reference's N2O table / `unique` output. */
size_t fgnn_k_ht_capacity(size_t max_items);
size_t fgnn_k_ht_bytes(size_t capacity);
/* Reset (cuda_hashtable.cu:714-723) */
int fgnn_k_ht_reset(void *table, size_t capacity, uint32_t *d_num_items,
                    fgnn_stream_t stream);
/* FillWithUnique (cuda_hashtable.cu:1017-1037): local = *d_num_items + index */
int fgnn_k_ht_fill_unique(void *table, size_t capacity, const uint32_t *input,
                          uint32_t n_max, const uint32_t *d_n, uint32_t *n2o,
                          uint32_t *d_num_items, fgnn_stream_t stream);
/* FillWithDuplicates (cuda_hashtable.cu:725-807): ids not yet present get
 * consecutive local ids in order of first occurrence in `input` (one legal
 * outcome of the reference's atomicCAS race, and exactly CPUHashTable0's order,
 * cpu_hashtable0.cc:37-47).  `pos` (n_max entries, scratch/out) receives the
 * bucket position of every input item so remapping needs no second probe. */
int fgnn_k_ht_fill_duplicates(void *table, size_t capacity,
                              const uint32_t *input, uint32_t n_max,
                              const uint32_t *d_n, uint32_t *pos,
                              uint32_t *n2o, uint32_t *d_num_items,
                              void *chain_ws, fgnn_stream_t stream);
/* GPUMapEdges for one column (cuda_mapping.cu:32-81).  With pos != NULL the
 * local id is read from bucket pos[i]; otherwise `global[i]` is probed. */
int fgnn_k_ht_map(const void *table, size_t capacity, const uint32_t *global,
                  const uint32_t *pos, uint32_t n_max, const uint32_t *d_n,
                  uint32_t *out_local, fgnn_stream_t stream);

/* FillWithDuplicates + GPUMapEdges in one pass (cuda_hashtable.cu:725-807 followed by
 * cuda_mapping.cu:68-81): like fgnn_k_ht_fill_duplicates, and out_local[i] receives
 * the local id of input[i] for every item of the fill. */
int fgnn_k_ht_fill_duplicates_map(void *table, size_t capacity,
                                  const uint32_t *input, uint32_t n_max,
                                  const uint32_t *d_n, uint32_t *pos,
                                  uint32_t *n2o, uint32_t *d_num_items,
                                  uint32_t *out_local, void *chain_ws,
                                  fgnn_stream_t stream);

/* ---- one whole mini-batch: DoGPUSample (cuda_loops.cc:50-267) ------------------- */
/* Enqueues hash-table reset, FillWithUnique(seeds) and, for layer i = L-1 .. 0,
 * sample -> FillWithDuplicates -> MapEdges on `stream` without any host round trip.
 * The uniform k-hop samplers insert their picks into the hash table while they
 * gather them and the remap is folded into the compaction pass, so a layer is two
 * launches (the reference: 8 kernels, 2 scans, ~12 stream syncs per layer).
 *   plan : what to sample and the per-stream scratch (reused by every batch of a slot)
 *   out  : where the batch's results go:
 *            n2o            running unique list; after the call n2o[0..num_src(0)) == input_nodes
 *            row[i]/col[i]  TrainGraph of layer i: row = neighbour local id, col = seed local id
 *                           (cuda_loops.cc:210-221); data[i] = visit counts (random walk only)
 *            counts         device u32[L][3] = {num_dst, num_edge, num_src} per layer
 * Layer i uses fanout[i]; in_max[i] bounds its inputs (PredictNumNodes, common.cc:330-339);
 * dst/pos/row/col/data[i] hold in_max[i]*fanout[i] entries. */
#define FGNN_MAX_LAYERS 8
typedef struct fgnn_sample_plan {
  int32_t sample_type; /* SampleType, common.h:50-58 */
  uint32_t num_layers;
  uint32_t fanout[FGNN_MAX_LAYERS];
  uint32_t in_max[FGNN_MAX_LAYERS];
  const uint32_t *indptr, *indices;
  const float *prob_table;
  const uint32_t *alias_table;
  const float *prob_prefix_table;
  uint32_t walk_len, num_walk; /* random walk (sample_type 3): fanout[i] = top-K */
  double restart_prob;
  uint64_t seed;
  void *table; /* fgnn_k_ht_bytes(capacity) */
  size_t capacity;
  uint32_t *num_items; /* device u32 */
  void *chain_ws;
  void *workspace; /* max over layers of the sampler's *_workspace_bytes, or NULL */
  size_t workspace_bytes;
  uint32_t *dst[FGNN_MAX_LAYERS]; /* scratch: sampled global ids */
  uint32_t *pos[FGNN_MAX_LAYERS]; /* scratch: their hash buckets */
  /* Versioned reset of the ordered hash table (cuda_hashtable.cu:714-723 keeps a `version` per bucket): 0 = the
   * call clears the table itself (memset); 1..126 = the caller guarantees that no bucket still carries this tag
   * (fgnn_k_ht_next_version hands out the tags and clears the table when they wrap), and the uniform k-hop
   * chain then never touches the buckets it does not use.  Honoured by sample_type khop2; other samplers clear. */
  uint32_t version;
  /* khop1 / weighted samplers: zero-initialised fgnn_k_seed_rank_workspace_bytes(num_nodes) bytes (or NULL: the
   * seeds are ordered by a library radix sort instead of the rank-by-bitmap) */
  uint32_t num_nodes;
  void *rank_ws;
} fgnn_sample_plan;
typedef struct fgnn_sample_out {
  uint32_t *n2o;
  uint32_t *row[FGNN_MAX_LAYERS], *col[FGNN_MAX_LAYERS], *data[FGNN_MAX_LAYERS];
  uint32_t *counts;
} fgnn_sample_out;
int fgnn_k_sample_batch(const fgnn_sample_plan *plan, const fgnn_sample_out *out,
                        const uint32_t *seeds, uint32_t n_seeds_max,
                        const uint32_t *d_n_seeds, uint64_t batch_key,
                        fgnn_stream_t stream);
/* "Super-batch": num_batches (<= FGNN_MAX_SUPER) independent mini-batches of the same configuration enqueued
 * together on ONE stream.  plans[k] / outs[k] are mini-batch k's own table, scratch and outputs (they may
 * differ in nothing else: topology, sampler, fanouts and capacity are taken from plans[0]).  For the uniform
 * k-hop sampler (khop2) every layer is two launches for ALL the mini-batches (gridDim.y = mini-batch), which
 * is what lets 8000-seed batches fill a 148-SM GPU; other samplers run batch after batch on the stream.
 * Results are exactly those of num_batches fgnn_k_sample_batch calls. */
#define FGNN_MAX_SUPER 8
int fgnn_k_sample_batch_multi(const fgnn_sample_plan *const *plans, const fgnn_sample_out *const *outs,
                              const uint32_t *const *seeds, const uint32_t *n_seeds_max,
                              const uint64_t *batch_keys, uint32_t num_batches,
                              fgnn_stream_t stream);
/* Next version tag of a table (host-side counter *state, start it at 0): returns 1..126 and clears the table
 * on `stream` whenever the tags wrap, so that plan->version = the returned tag is always safe. */
uint32_t fgnn_k_ht_next_version(uint32_t *state, void *table, size_t capacity, fgnn_stream_t stream);

/* ---- feature cache ---------------------------------------------------------- */
/* SampleCacheTableInit / DistCacheManager ctor steps 1-2 (dist_engine.cc:193-229,
 * dist_cache_manager_host.cc:84-95): table[v]=EMPTY; table[rank[i]]=i, i<num_cached */
int fgnn_k_cache_table_build(uint32_t *table, size_t num_nodes,
                             const uint32_t *ranking_nodes, size_t num_cached,
                             fgnn_stream_t stream);
/* GetMissCacheIndex (cuda_cache.cu:162-234): stable split of `nodes`.
 * d_counts[0] <- num_miss, d_counts[1] <- num_cache. */
int fgnn_k_cache_split(const uint32_t *table, const uint32_t *nodes,
                       uint32_t n_max, const uint32_t *d_n, uint32_t *miss_src,
                       uint32_t *miss_dst, uint32_t *cache_src,
                       uint32_t *cache_dst, uint32_t *d_counts, void *chain_ws,
                       fgnn_stream_t stream);

/* ---- extraction --------------------------------------------------------------- */
/* Generic row copy dst[dst_index?[i]] = src[src_index?[i] & src_mask] with rows
 * of row_bytes.  One kernel covers GPUExtract (cuda_extraction.cu:74-117),
 * combine_miss_data and combine_cache_data (dist_cache_manager_device.cu:37-82)
 * and, with `src` in pinned host memory, extract_miss_data
 * (dist_cache_manager_host.cc:38-56).  src_mask = ~0 normally; (1<<k)-1 for
 * SAMGRAPH_EMPTY_FEAT=k (cuda_extraction.cu:131). */
int fgnn_k_row_copy(void *dst, const uint32_t *dst_index, const void *src,
                    const uint32_t *src_index, uint64_t src_mask,
                    uint32_t n_max, const uint32_t *d_n, size_t row_bytes,
                    fgnn_stream_t stream);

/* Fused cache-aware gather, replaces GetMissCacheIndex + ExtractMissData +
 * H2D + CombineMissData + CombineCacheData (dist_loops.cc:713-846):
 *   slot = table[nodes[i]];
 *   out[i] = slot != EMPTY ? shard[slot % num_shards][slot / num_shards]
 *                          : miss_src[nodes[i] & miss_mask]
 * `shards` is a DEVICE array of num_shards row-major cache shard base pointers
 * (local HBM or NVLink peer mappings); num_shards == 1 is the replicated cache
 * of the reference.  `miss_src` may be pinned host memory (UVA) or device
 * memory.  d_stats (optional, 2 x u64): += {hit rows, miss rows}. */
int fgnn_k_gather_cached(void *out, const uint32_t *nodes, uint32_t n_max,
                         const uint32_t *d_n, const uint32_t *table,
                         const void *const *shards, uint32_t num_shards,
                         const void *miss_src, uint64_t miss_mask,
                         size_t row_bytes, unsigned long long *d_stats,
                         fgnn_stream_t stream);
/* The same gather over a HYBRID cache layout (BASELINE north_star: cache partitioned across the trainer GPUs and
 * served by NVLink peer loads): the hottest `num_replicated` slots are held by every trainer (`replica`, local
 * HBM), the remaining slots are striped over the trainers' shards: slot s >= R lives on shard (s-R) % num_shards
 * at local row (s-R) / num_shards.  With a power-law hotness ranking a few % of replicated rows take most of the
 * hits off the NVLink path.  num_replicated = 0 is the plain striped layout of fgnn_k_gather_cached.
 * d_stats[0] += hits, d_stats[1] += misses; *d_remote += rows read from a peer's shard (optional). */
typedef struct fgnn_cache_layout {
  const uint32_t *table;      /* node -> slot, EMPTY = not cached */
  const void *const *shards;  /* DEVICE array of num_shards shard base pointers */
  uint32_t num_shards, self_shard;
  const void *replica;        /* rows of slots [0, num_replicated) on this GPU, or NULL */
  uint32_t num_replicated;
  const void *miss_src;       /* pinned host feature table (UVA) */
  uint64_t miss_mask;
  size_t row_bytes;
  /* Optional second pass for the rows of PEER stripes: with `defer_ws` (fgnn_k_gather_defer_workspace_bytes(n_max)
   * bytes of device memory, zero before the first call) the main kernel does not wait for peer rows: it lists
   * them, and a second launch on the same stream copies all listed rows at once, one warp per row.  Measured on
   * 4 x B200: a peer row costs a warp of the pipelined ring 30-40 us when every GPU's HBM is saturated by its own
   * gather, so a few % of peer rows tripled the kernel time (profiles/r2_partition_diag_n4.txt).  NULL = peer rows
   * inside the ring (round 1 behaviour). */
  void *defer_ws;
} fgnn_cache_layout;
size_t fgnn_k_gather_defer_workspace_bytes(uint32_t n_max);
int fgnn_k_gather_cached_layout(void *out, const uint32_t *nodes, uint32_t n_max, const uint32_t *d_n,
                                const fgnn_cache_layout *layout, unsigned long long *d_stats,
                                unsigned long long *d_remote, fgnn_stream_t stream);


/* ---- partitioned cache plumbing ------------------------------------------------ */
/* The reference replicates the feature cache in every trainer process
 * (dist_engine.cc:418-424).  Here each trainer GPU owns the slots with
 * slot % T == t; the other trainers map that shard (same process: peer access,
 * other processes: CUDA IPC) and fgnn_k_gather_cached reads it over NVLink.
 * shard_alloc returns plain cudaMalloc memory (exportable); the 64-byte handle is
 * a cudaIpcMemHandle_t. */
#define FGNN_IPC_HANDLE_BYTES 64
int fgnn_k_shard_alloc(void **ptr, size_t bytes);
int fgnn_k_shard_free(void *ptr);
int fgnn_k_ipc_export(void *ptr, void *handle64);
int fgnn_k_ipc_open(const void *handle64, void **ptr);
int fgnn_k_ipc_close(void *ptr);
int fgnn_k_enable_peer(int peer_device);

/* ---- PreSC ---------------------------------------------------------------------- */
/* freq[nodes[i]] += 1  (cuda/pre_sampler.cc:84-88) */
int fgnn_k_freq_count(uint32_t *freq, const uint32_t *nodes, uint32_t n_max,
                      const uint32_t *d_n, fgnn_stream_t stream);
/* ranking_nodes = argsort desc of u64 {freq:hi32,id:lo32}
 * (cuda/pre_sampler.cc:44-49,97-99,121-142) */
size_t fgnn_k_presc_rank_workspace_bytes(size_t num_nodes);
int fgnn_k_presc_rank(const uint32_t *freq, size_t num_nodes,
                      uint32_t *ranking_nodes, void *workspace,
                      size_t workspace_bytes, fgnn_stream_t stream);

/* ---- dataset preparation (reference: offline CPU tools) --------------- */
/* prob_table / alias_table of the alias sampler from per-edge weights: Vose's method with two FIFO
 * queues per CSR row, bit-identical to utility/data-process/toolkit/weight/create_alias_table.cc:96-180
 * for the same weights (alias = neighbour id).  workspace: fgnn_k_alias_table_workspace_bytes(E).
 * Not re-entrant across streams (one global row ticket). */
size_t fgnn_k_alias_table_workspace_bytes(size_t num_edges);
int fgnn_k_build_alias_table(const uint32_t *indptr, const uint32_t *indices,
                             size_t num_nodes, size_t num_edges, const float *weights,
                             float *prob_table, uint32_t *alias_table, void *workspace,
                             size_t workspace_bytes, fgnn_stream_t stream);
/* prob_prefix_table[off+i] = sequential fp32 sum of the row's weights up to i
 * (toolkit/weight/create_prob_prefix_table.cc:94-123) */
int fgnn_k_build_prefix_table(const uint32_t *indptr, size_t num_nodes, const float *weights,
                              float *prob_prefix_table, fgnn_stream_t stream);
/* out_degree[v] = occurrences of v in indices (common/graph_loader.cc:109-147) */
int fgnn_k_out_degree(const uint32_t *indices, size_t num_edges, uint32_t *out_degree,
                      size_t num_nodes, fgnn_stream_t stream);
/* cache_by_degree ranking: sort {out_degree, id} descending (toolkit/cache/cache_by_degree.cc:36-58).
 * out_degree u32[V] is filled as a by-product; workspace: fgnn_k_presc_rank_workspace_bytes(V). */
int fgnn_k_rank_by_degree(const uint32_t *indices, size_t num_edges, size_t num_nodes,
                          uint32_t *out_degree, uint32_t *ranking_nodes, void *workspace,
                          size_t workspace_bytes, fgnn_stream_t stream);
/* cache_by_random ranking: a seeded uniform permutation of [0, V) (toolkit/cache/cache_by_random.cc:36-48;
 * Philox-key sort instead of the reference's default-seeded mt19937 Fisher-Yates). */
size_t fgnn_k_rank_random_workspace_bytes(size_t num_nodes);
int fgnn_k_rank_random(size_t num_nodes, uint64_t seed, uint32_t *ranking_nodes, void *workspace,
                       size_t workspace_bytes, fgnn_stream_t stream);

/* cache_by_heuristic ranking (toolkit/cache/cache_by_heuristic.cc:28-91): training nodes, then their first-hop
 * neighbours in order of first appearance, then everything else by {out_degree, id} descending.  Built from the
 * sampler's own ordered hash table + the cache_by_degree ranking + the stable split of fgnn_k_cache_split.
 * num_neighbours = sum of the train nodes' row lengths (fgnn_k_row_len_sum writes it to *d_total on the device;
 * the caller reads it back once to size the workspace); train_set must not contain duplicates. */
int fgnn_k_row_len_sum(const uint32_t *indptr, const uint32_t *nodes, size_t n,
                       unsigned long long *d_total, fgnn_stream_t stream);
size_t fgnn_k_rank_heuristic_workspace_bytes(size_t num_nodes, size_t num_train, size_t num_neighbours);
int fgnn_k_rank_by_heuristic(const uint32_t *indptr, const uint32_t *indices, size_t num_nodes,
                             size_t num_edges, const uint32_t *train_set, size_t num_train,
                             size_t num_neighbours, uint32_t *ranking_nodes, void *workspace,
                             size_t workspace_bytes, fgnn_stream_t stream);

/* ---- launch event trace (profiling aid, off by default) --------------------------------------- */
/* fgnn_k_trace_enable(1, max_records): from now on fgnn_k_sample_batch and fgnn_k_gather_cached record a CUDA
 * event after each of their launches (time zero = an event on the legacy default stream at enable time).
 * fgnn_k_trace_dump waits for the recorded events, writes (label, stream handle, milliseconds since time zero)
 * for up to max_records of them, switches the trace off and returns the number written.  A record marks the END
 * of the named launch on its stream; the launch began at the previous record of the same stream (or later, when
 * the GPU was busy elsewhere).  Not thread-safe against concurrent launches from several host threads beyond a
 * mutex around the record list. */
#define FGNN_TRACE_BATCH_BEGIN 0
#define FGNN_TRACE_TABLE_RESET 1
#define FGNN_TRACE_FILL_SEEDS 2
#define FGNN_TRACE_LAYER(i) (100 * ((i) + 1))
#define FGNN_TRACE_SAMPLE 1
#define FGNN_TRACE_INSERT 2
#define FGNN_TRACE_COMPACT 3
#define FGNN_TRACE_MAP 4
#define FGNN_TRACE_GATHER_BEGIN 900
#define FGNN_TRACE_GATHER_END 901
int fgnn_k_trace_enable(int on, size_t max_records);
long fgnn_k_trace_dump(size_t max_records, int *labels, uint64_t *streams, float *ms);

/* ---- block hand-off in CSC form ------------------------------------------ */
/* (row, col) COO of one sampled layer -> the three arrays of the reference's DGL patch
 * `create_unitgraph_from_csc` (3rdparty/dgl.patch:30-57), replacing DGL's own COO->CSC conversion of the
 * block built at samgraph/torch/adapter.py:92-95 (the stage timed as kLogL1ConvertTime):
 *   indptr   u32[num_dst+1]  indptr[d] = #edges with col < d
 *   indices  u32[e]          row ids ordered by (col, edge id)   (stable)
 *   edge_ids u32[e] or NULL  the permutation (original edge index of every CSC entry)
 * col_sorted != 0: col is non-decreasing (khop0/khop2/hash-dedup/random-walk blocks are seed-major), the
 * permutation is the identity, no workspace is needed and `indices` may be NULL or == row.
 * col_sorted == 0: one radix sort over log2(num_dst) key bits; workspace from
 * fgnn_k_coo_to_csc_workspace_bytes(e_max, num_dst).  *d_e (may be NULL = e_max) is read on the device. */
size_t fgnn_k_coo_to_csc_workspace_bytes(uint32_t e_max, uint32_t num_dst);
int fgnn_k_coo_to_csc(const uint32_t *row, const uint32_t *col, uint32_t e_max, const uint32_t *d_e,
                      uint32_t num_dst, int col_sorted, uint32_t *indptr, uint32_t *indices,
                      uint32_t *edge_ids, void *workspace, size_t workspace_bytes, fgnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FGNN_KERNELS_H */
