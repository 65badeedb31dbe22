/* fgnn_dataset_tools.h — host-side builders of the two cache rankings the reference computes with OFFLINE host
 * tools and loads from files (engine.cc:233-244: cache_by_degree_hop.bin, cache_by_fake_optimal.bin).  They are
 * dataset preparation, not part of the per-batch path: integer / fp64 host code like the reference's tools, run
 * once (tools/make_dataset.py --cache-policy-files, or by the engine at data_init when the file is absent).
 * The GPU builders of the other rankings (degree, heuristic, random, PreSC) are in fgnn_kernels.h.
 *
 * Both write `ranking_nodes u32[num_nodes]` (hottest first), the format of the reference's cache_by_*.bin files,
 * and are bit-exact against the reference tools' output (tests/golden/ref_cache_policy_golden.npz).
 * Return 0 on success, -1 on invalid arguments.
 */
#ifndef FGNN_DATASET_TOOLS_H
#define FGNN_DATASET_TOOLS_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* utility/data-process/toolkit/cache/cache_by_degree_hop.cc:31-180: vertices within `hops` (the tool: 2) hops of
 * the training set come first, ordered by their out-degree inside the sub-graph made of those vertices' adjacency
 * rows; every other vertex follows by out-degree in the whole graph.  Ties: larger id first (std::greater on
 * {degree, id} pairs).  `num_threads` = worker threads (0 = all cores); the result does not depend on it. */
int fgnn_rt_rank_degree_hop(const uint32_t *indptr, const uint32_t *indices, size_t num_nodes,
                            const uint32_t *train_set, size_t num_train, int hops, int num_threads,
                            uint32_t *ranking_nodes);

/* utility/data-process/toolkit/cache/cache_by_fake_optimal.cc:61-183: expected number of training nodes whose
 * 2-hop sample (fanout {fanout0, fanout1}; the tool hard-codes {25, 10}, the hop next to the seed uses fanout1)
 * touches each vertex, one training node at a time; vertices sorted by that expectation, descending, ties larger
 * id first.  The tool's floating-point products are taken in an order that depends on its -t option (vertices are
 * bucketed by id % threads, cache_by_fake_optimal.cc:44-60); `order_threads` reproduces that order (the tool's
 * default is 48, options.cc:28) and has nothing to do with `num_threads`, the worker threads used here. */
int fgnn_rt_rank_fake_optimal(const uint32_t *indptr, const uint32_t *indices, size_t num_nodes,
                              const uint32_t *train_set, size_t num_train, int fanout0, int fanout1,
                              int order_threads, int num_threads, uint32_t *ranking_nodes);

#ifdef __cplusplus
}
#endif
#endif
