/* TEST INFRASTRUCTURE — CPU restatement ("oracle") of the reference's
 * sampling-and-extraction hot path.  Never imported by the product path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
 * may load it.
 *
 * Parity pins:
 *   - unique/remap, extraction, khop0/khop2 copy path and output layout are
 *     pinned against the reference's own CPU translation units compiled from
 *     /root/reference into oracle/_ref (tests/test_oracle_vs_ref.py) and the
 *     fixtures generated from them (tests/golden/).
 *   - the reference has NO golden vectors and seeds cuRAND XORWOW from the wall
 *     clock (cuda_random_states.cu:105-107), so RNG-dependent outputs are
 *     "parity unpinned" at the bit level: the RNG is replaced by the
 *     counter-based Philox4x32-10 defined here (checked against the published
 *     Random123 known-answer vectors) and compared with the reference
 *     distributionally.
 *
 * All "file:line" citations are relative to /root/reference/samgraph/common/.
 */
#ifndef FGNN_ORACLE_H
#define FGNN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGO_EMPTY 0xFFFFFFFFu /* Constant::kEmptyKey, constant.h:71 */

/* ---- RNG ---------------------------------------------------------------- */
void fgo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]);
/* draw #`draw` of item `item` in stream (seed, batch_key, tag). */
uint32_t fgo_rand_u32(uint64_t seed, uint64_t batch_key, uint32_t tag,
                      uint32_t item, uint32_t draw);
float fgo_uniform_f32(uint32_t x);               /* curand_uniform.h:69-72 */
double fgo_uniform_f64(uint32_t x, uint32_t y);  /* curand_uniform.h:101-106 */

/* ---- sizing ------------------------------------------------------------- */
size_t fgo_predict_num_nodes(size_t batch, const size_t *fanout,
                             size_t num_fanout); /* common.cc:330-339 */
size_t fgo_table_size(size_t num, size_t scale); /* cuda_hashtable.cu:125-128 */

/* ---- epoch shuffle (replaces cuda_shuffler.cc:75-126 / dist_shuffler.cc:98-151):
 * stable ascending sort of train_set by key_i = rand(seed, epoch, FGO_SHUFFLE_TAG, i, 0) */
#define FGO_SHUFFLE_TAG 0xFFFF0001u
void fgo_shuffle(const uint32_t *train_set, size_t n, uint64_t seed,
                 uint64_t epoch, uint32_t *out);

/* ---- samplers: all write compact COO (out_src = seed id, out_dst = nbr id)
 * and return the number of edges.  out_* must hold num_input*fanout. -------- */
/* cuda_sampling_khop0.cu:42-89 + :128-173 (reservoir, Algorithm R) */
size_t fgo_sample_khop0(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst);
/* cuda_sampling_khop2.cu:42-87 (partial Fisher-Yates), stateless: the
 * adjacency list is treated as immutable, swaps happen in a virtual copy. */
size_t fgo_sample_khop2(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst);
/* cuda_sampling_khop1.cu:42-127,169-178: with replacement, stable sort by
 * src id, drop entries equal to their successor. */
size_t fgo_sample_khop1(const uint32_t *indptr, const uint32_t *indices,
                        const uint32_t *input, size_t num_input, size_t fanout,
                        uint64_t seed, uint64_t batch_key, uint32_t tag,
                        uint32_t *out_src, uint32_t *out_dst);
/* cuda_sampling_weighted_khop.cu:41-128,165-178 (alias method) */
size_t fgo_sample_weighted_khop(const uint32_t *indptr, const uint32_t *indices,
                                const float *prob_table,
                                const uint32_t *alias_table,
                                const uint32_t *input, size_t num_input,
                                size_t fanout, uint64_t seed,
                                uint64_t batch_key, uint32_t tag,
                                uint32_t *out_src, uint32_t *out_dst);
/* cuda_sampling_weighted_khop_prefix.cu:41-92 */
size_t fgo_sample_weighted_khop_prefix(
    const uint32_t *indptr, const uint32_t *indices,
    const float *prob_prefix_table, const uint32_t *input, size_t num_input,
    size_t fanout, uint64_t seed, uint64_t batch_key, uint32_t tag,
    uint32_t *out_src, uint32_t *out_dst);
/* cuda_sampling_weighted_khop_hash_dedup.cu:41-119; max_draws bounds the
 * rejection loop (the reference loops forever on a multigraph row). */
size_t fgo_sample_weighted_khop_hash_dedup(
    const uint32_t *indptr, const uint32_t *indices, const float *prob_table,
    const uint32_t *alias_table, const uint32_t *input, size_t num_input,
    size_t fanout, uint64_t seed, uint64_t batch_key, uint32_t tag,
    uint32_t *out_src, uint32_t *out_dst);

/* cuda_sampling_random_walk.cu:43-109: fills tmp_src/tmp_dst[num_input*W*L]
 * (tmp_src == FGO_EMPTY marks a dead step). */
void fgo_random_walk(const uint32_t *indptr, const uint32_t *indices,
                     const uint32_t *input, size_t num_input, size_t walk_len,
                     double restart_prob, size_t num_walk, uint64_t seed,
                     uint64_t batch_key, uint32_t tag, uint32_t *tmp_src,
                     uint32_t *tmp_dst);
/* cuda_frequency_hashmap.cu:361-401,460-507,585-607,644-676,1143-1367:
 * per start node, count visit multiplicities, order by (count desc, first
 * occurrence asc), keep K.  out_* must hold num_input*K.  Returns #edges. */
size_t fgo_topk(const uint32_t *tmp_src, const uint32_t *tmp_dst,
                const uint32_t *input, size_t num_input, size_t edges_per_node,
                size_t K, uint32_t *out_src, uint32_t *out_dst,
                uint32_t *out_data);

/* ---- ordered hash table (cuda_hashtable.cu:131-174,387-438,725-807,
 * 1017-1037; canonical order = first occurrence = cpu_hashtable0.cc:37-47) -- */
typedef struct fgo_hashtable fgo_hashtable;
fgo_hashtable *fgo_hashtable_new(size_t max_items);
void fgo_hashtable_free(fgo_hashtable *t);
void fgo_hashtable_reset(fgo_hashtable *t);
size_t fgo_hashtable_num_items(const fgo_hashtable *t);
void fgo_hashtable_fill_unique(fgo_hashtable *t, const uint32_t *input,
                               size_t n);
void fgo_hashtable_fill_duplicates(fgo_hashtable *t, const uint32_t *input,
                                   size_t n);
/* copies the first n entries of the new->old list */
void fgo_hashtable_unique(const fgo_hashtable *t, uint32_t *out, size_t n);
/* cuda_mapping.cu:32-81; returns -1 if an id is absent */
int fgo_hashtable_map(const fgo_hashtable *t, const uint32_t *in, size_t n,
                      uint32_t *out);

/* ---- cache ---------------------------------------------------------------- */
/* dist_engine.cc:193-229 / dist_cache_manager_host.cc:84-95:
 * table[v] = FGO_EMPTY; table[rank[i]] = i for i < num_cached */
void fgo_cache_table_build(const uint32_t *ranking_nodes, size_t num_nodes,
                           size_t num_cached, uint32_t *table);
size_t fgo_num_cached(size_t num_nodes, double cache_percentage);
/* cuda_cache.cu:33-158: stable split */
void fgo_cache_split(const uint32_t *table, const uint32_t *nodes, size_t n,
                     uint32_t *miss_src, uint32_t *miss_dst, size_t *num_miss,
                     uint32_t *cache_src, uint32_t *cache_dst,
                     size_t *num_cache);
/* generic row copy: dst[dst_index?[i]] = src[(src_index?[i]) & mask]; covers
 * GPUExtract (cuda_extraction.cu:31-49), extract_miss_data
 * (dist_cache_manager_host.cc:38-56), combine_miss_data / combine_cache_data
 * (dist_cache_manager_device.cu:37-82). */
void fgo_row_copy(void *dst, const uint32_t *dst_index, const void *src,
                  const uint32_t *src_index, size_t n, size_t row_bytes,
                  uint64_t src_index_mask);

/* ---- PreSC (cuda/pre_sampler.cc:57-110,121-142) ---------------------------- */
void fgo_freq_count(uint32_t *freq, const uint32_t *nodes, size_t n);
/* rank by u64 {freq:hi32,id:lo32} descending */
void fgo_presc_rank(const uint32_t *freq, size_t num_nodes, uint32_t *rank);

/* ---- weight tables (utility/data-process/toolkit/weight/) ------------------ */
/* create_alias_table.cc:96-180 with caller-supplied per-edge weights */
void fgo_build_alias_table(const uint32_t *indptr, const uint32_t *indices,
                           size_t num_nodes, const float *weights,
                           float *prob_table, uint32_t *alias_table);
/* create_prob_prefix_table.cc:94-123 */
void fgo_build_prefix_table(const uint32_t *indptr, size_t num_nodes,
                            const float *weights, float *prefix_table);

void fgo_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif /* FGNN_ORACLE_H */
