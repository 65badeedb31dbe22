// fgnn_k_sample_batch: one mini-batch of DoGPUSample (cuda_loops.cc:50-267 ==
// dist_loops.cc:51-269) enqueued on one stream with no host round trip.  This is
// the single place where the per-layer kernel sequence lives; the C++ engine
// (csrc/runtime) and the Python driver (fgnn_b200/pipeline.py) both call it.
#include "hashtable.cuh"

#include <stdlib.h>

using namespace fgnn;

extern "C" int fgnn_k_sample_batch(const fgnn_sample_plan *pl, const fgnn_sample_out *out,
                                   const uint32_t *seeds, uint32_t n_seeds_max,
                                   const uint32_t *d_n_seeds, uint64_t batch_key,
                                   fgnn_stream_t stream) {
  if (!pl || !out || !pl->indptr || !pl->indices || !pl->table || !pl->num_items || !pl->chain_ws ||
      !out->n2o || !out->counts)
    return FGNN_ERR_BAD_ARG;
  const uint32_t L = pl->num_layers;
  if (L == 0 || L > FGNN_MAX_LAYERS) return FGNN_ERR_UNSUPPORTED;
  if ((pl->capacity & (pl->capacity - 1)) || pl->capacity > 0x80000000ull) return FGNN_ERR_BAD_ARG;
  if (n_seeds_max > pl->in_max[L - 1]) return FGNN_ERR_BAD_ARG;
  if (n_seeds_max > 0 && !seeds) return FGNN_ERR_BAD_ARG;
  for (uint32_t i = 0; i < L; ++i) {
    if (!pl->dst[i] || !pl->pos[i] || !out->row[i] || !out->col[i]) return FGNN_ERR_BAD_ARG;
    if ((uint64_t)pl->in_max[i] * pl->fanout[i] >= 0x7FFFFFFFull) return FGNN_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t *counts = out->counts;
  // A/B switch for profiling: bit 0 = samplers insert their picks while gathering them (measured SLOWER on
  // B200, r1_n: 177 vs 145 us per batch — the smem-limited sampler CTAs have too few threads to hide the
  // table's L2 latency — so it is off), bit 1 = remap folded into the compaction pass (on)
  // bit 2 = khop2 writes a padded [seed][fanout] block and ONE chained scan (in the unique/remap pass) compacts
  // the edges and numbers the new ids (3 launches per layer, one look-back chain instead of two; ncu r1_q c4/c7:
  // sampler 18+39 us -> 10+24 us per batch; on)
  // bit 3 = the two-launch-per-layer chain of fast_chain.cu (sampler with compacted Fisher-Yates lanes that
  // inserts its picks, register-resident compaction with a direct-sum scan, versioned table; on)
  const char *fz = getenv("FGNN_BATCH_FUSE");
  const int fuse = fz && *fz ? atoi(fz) : 14;
  if ((fuse & 8) && fast_chain_supported(pl)) {
    trace_mark(st, FGNN_TRACE_BATCH_BEGIN);
    return fast_chain_launch(&pl, &out, &seeds, &n_seeds_max, &d_n_seeds, &batch_key, 1, st);
  }

  // Reset (cuda_loops.cc:63) + FillWithUnique of the seeds (:67-69); the seed count doubles as the
  // input count of the first sampled layer
  trace_mark(st, FGNN_TRACE_BATCH_BEGIN);
  cudaError_t e = cudaMemsetAsync(pl->table, 0xFF, fgnn_k_ht_bytes(pl->capacity), st);
  if (e != cudaSuccess) return (int)e;
  trace_mark(st, FGNN_TRACE_TABLE_RESET);
  int rc = ht_fill_unique_first_launch(pl->table, pl->capacity, seeds, n_seeds_max, d_n_seeds, out->n2o,
                                       pl->num_items, counts + 3 * (L - 1), st);
  if (rc) return rc;
  trace_mark(st, FGNN_TRACE_FILL_SEEDS);

  for (int i = (int)L - 1; i >= 0; --i) {  // cuda_loops.cc:87
    uint32_t *n_in = counts + 3 * i, *n_edge = counts + 3 * i + 1, *n_src = counts + 3 * i + 2;
    uint32_t *n_in_next = i > 0 ? counts + 3 * (i - 1) : nullptr;
    const uint32_t nmax = pl->in_max[i], f = pl->fanout[i];
    const uint32_t emax = nmax * f;
    fgnn_rng rng{pl->seed, batch_key, (uint32_t)i};
    uint32_t *dst = pl->dst[i], *col = out->col[i];
    bool inserted = false;
    if (pl->sample_type == 5 && (fuse & 4)) {
      rc = sample_khop2_pad_launch(pl->indptr, pl->indices, out->n2o, nmax, n_in, f, rng, dst, st);
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_SAMPLE);
      rc = ht_insert_launch(pl->table, pl->capacity, dst, emax, n_in, pl->pos[i], st, f);
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_INSERT);
      rc = ht_compact_pad_launch(pl->table, dst, nmax, n_in, f, pl->pos[i], out->n2o, pl->num_items, out->row[i],
                                 col, n_edge, n_src, n_in_next, pl->chain_ws, st);
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_COMPACT);
      continue;
    }
    switch (pl->sample_type) {  // cuda_loops.cc:118-161
      case 0:
      case 5: {
        HtInsert ht{(fuse & 1) ? (Bucket *)pl->table : nullptr, (uint32_t)(pl->capacity - 1), pl->pos[i]};
        rc = sample_khop_launch(pl->sample_type == 0 ? 0 : 2, pl->indptr, pl->indices, out->n2o, nmax, n_in,
                                f, rng, nullptr, dst, col, n_edge, pl->chain_ws, ht, st);
        inserted = (fuse & 1) != 0;
        break;
      }
      case 1:
      case 2:
      case 4:
        rc = fgnn_k_sample_replace_ranked(pl->sample_type, pl->indptr, pl->indices, pl->prob_table, pl->alias_table,
                                          pl->prob_prefix_table, out->n2o, nmax, n_in, f, rng, nullptr, dst, col,
                                          n_edge, pl->workspace, pl->workspace_bytes, pl->chain_ws, pl->rank_ws,
                                          pl->num_nodes, stream);
        break;
      case 6:
        rc = fgnn_k_sample_weighted_hash_dedup(pl->indptr, pl->indices, pl->prob_table, pl->alias_table,
                                               out->n2o, nmax, n_in, f, rng, nullptr, dst, col, n_edge,
                                               pl->chain_ws, stream);
        break;
      case 3:
        if (!out->data[i]) return FGNN_ERR_BAD_ARG;
        rc = fgnn_k_sample_random_walk(pl->indptr, pl->indices, out->n2o, nmax, n_in, pl->walk_len,
                                       pl->restart_prob, pl->num_walk, f, rng, nullptr, dst, col,
                                       out->data[i], n_edge, nullptr, nullptr, pl->workspace,
                                       pl->workspace_bytes, pl->chain_ws, stream);
        break;
      default:
        return FGNN_ERR_BAD_ARG;
    }
    if (rc) return rc;
    trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_SAMPLE);
    // populate the hash table with the sampled neighbours and remap them (:176-205)
    if (!inserted) {
      rc = ht_insert_launch(pl->table, pl->capacity, dst, emax, n_edge, pl->pos[i], st);
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_INSERT);
    }
    rc = ht_compact_launch(pl->table, pl->capacity, dst, emax, n_edge, pl->pos[i], out->n2o, pl->num_items,
                           (fuse & 2) ? out->row[i] : nullptr, n_src, n_in_next, pl->chain_ws, st);
    if (rc) return rc;
    trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_COMPACT);
    if (!(fuse & 2)) {
      rc = fgnn_k_ht_map(pl->table, pl->capacity, nullptr, pl->pos[i], emax, n_edge, out->row[i], stream);
      if (rc) return rc;
      trace_mark(st, FGNN_TRACE_LAYER(i) + FGNN_TRACE_MAP);
    }
  }
  return 0;
}

extern "C" int fgnn_k_sample_batch_multi(const fgnn_sample_plan *const *plans, const fgnn_sample_out *const *outs,
                                         const uint32_t *const *seeds, const uint32_t *n_seeds_max,
                                         const uint64_t *batch_keys, uint32_t num_batches,
                                         fgnn_stream_t stream) {
  if (!plans || !outs || !seeds || !n_seeds_max || !batch_keys) return FGNN_ERR_BAD_ARG;
  if (num_batches == 0) return 0;
  if (num_batches > FGNN_MAX_SUPER) return FGNN_ERR_UNSUPPORTED;
  const char *fz = getenv("FGNN_BATCH_FUSE");
  const int fuse = fz && *fz ? atoi(fz) : 14;
  bool fast = (fuse & 8) != 0;
  for (uint32_t k = 0; k < num_batches; ++k) {
    const fgnn_sample_plan *pl = plans[k];
    if (!pl || !outs[k] || !pl->indptr || !pl->indices || !pl->table || !pl->num_items || !pl->chain_ws ||
        !outs[k]->n2o || !outs[k]->counts)
      return FGNN_ERR_BAD_ARG;
    const uint32_t L = pl->num_layers;
    if (L == 0 || L > FGNN_MAX_LAYERS) return FGNN_ERR_UNSUPPORTED;
    if ((pl->capacity & (pl->capacity - 1)) || pl->capacity > 0x80000000ull) return FGNN_ERR_BAD_ARG;
    if (n_seeds_max[k] > pl->in_max[L - 1] || (n_seeds_max[k] > 0 && !seeds[k])) return FGNN_ERR_BAD_ARG;
    // one configuration per super-batch
    if (pl->sample_type != plans[0]->sample_type || L != plans[0]->num_layers || pl->capacity != plans[0]->capacity ||
        pl->indptr != plans[0]->indptr || pl->indices != plans[0]->indices || pl->seed != plans[0]->seed)
      return FGNN_ERR_BAD_ARG;
    for (uint32_t i = 0; i < L; ++i) {
      if (pl->fanout[i] != plans[0]->fanout[i] || pl->in_max[i] != plans[0]->in_max[i]) return FGNN_ERR_BAD_ARG;
      if (!pl->pos[i] || !outs[k]->row[i] || !outs[k]->col[i]) return FGNN_ERR_BAD_ARG;
    }
    for (uint32_t j = 0; j < k; ++j)
      if (plans[j]->table == pl->table || outs[j]->n2o == outs[k]->n2o || plans[j]->chain_ws == pl->chain_ws)
        return FGNN_ERR_BAD_ARG;  // mini-batches of one launch run concurrently: nothing may be shared
    if (!fast_chain_supported(pl)) fast = false;
  }
  if (fast) {
    trace_mark((cudaStream_t)stream, FGNN_TRACE_BATCH_BEGIN);
    return fast_chain_launch(plans, outs, seeds, n_seeds_max, nullptr, batch_keys, num_batches, (cudaStream_t)stream);
  }
  for (uint32_t k = 0; k < num_batches; ++k) {
    const int rc = fgnn_k_sample_batch(plans[k], outs[k], seeds[k], n_seeds_max[k], nullptr, batch_keys[k], stream);
    if (rc) return rc;
  }
  return 0;
}

extern "C" uint32_t fgnn_k_ht_next_version(uint32_t *state, void *table, size_t capacity, fgnn_stream_t stream) {
  if (!state || !table) return 0;
  uint32_t v = *state + 1;
  if (*state == 0 || v > 126) {  // first use or wrap: no bucket may keep a tag that is about to be reused
    cudaMemsetAsync(table, 0xFF, capacity * sizeof(Bucket), (cudaStream_t)stream);
    v = 1;
  }
  *state = v;
  return v;
}
