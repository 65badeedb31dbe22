"""BASELINE config #1: GCN 3-layer, fanout list [5,10,15] (sampled 15 -> 10 -> 5), batch 8000, on an
ogbn-products-shaped synthetic graph with SamGraph's CPU sampler / extractor — the one configuration the reference
runs without a GPU.  The reference's OWN code (oracle/_ref: CPUSampleKHop2, CPUHashTable2 Populate / MapNodes /
MapEdges, CPUExtract, driven like DoCPUSample, cpu_loops.cc:55-191) produces every mini-batch; the oracle is fed
the reference's sampled edges (the CPU sampler draws from a thread-local mt19937, cpu_random.cc:26-30, so the draws
themselves cannot be replayed) and must reproduce unique list, remapped blocks and extracted rows bit for bit —
parity definition (2) of SURVEY §8c."""
import numpy as np
import pytest


def gcn_batches(V, T, batch, rng):
    train = rng.permutation(V)[:T].astype(np.uint32)
    for lo in range(0, T, batch):
        yield train[lo:lo + batch]


@pytest.mark.parametrize("scale,num_batches", [(16, 3), (64, 2)])
def test_gcn_products_cpu_pipeline_reference_vs_oracle(oracle, ref, scale, num_batches):
    from fgnn_b200.synth import SHAPES, make_graph_numpy
    V, E, D, C, T = SHAPES["products"]
    V, E, T = V // scale, E // scale, max(8000, T // scale)
    fanout = [5, 10, 15]
    indptr, indices = make_graph_numpy(V, E, seed=31)
    rng = np.random.default_rng(scale)
    feat = (rng.random((V, D), dtype=np.float32) * 2 - 1).astype(np.float32)
    label = rng.integers(0, C, size=V).astype(np.int64)
    ref.set_threads(1)                      # 1 thread: CPUHashTable2's CAS order is the input order
    assert ref.predict_num_nodes(8000, fanout) == oracle.predict_num_nodes(8000, fanout) == 8448000
    deg = np.diff(indptr.astype(np.int64))
    for b, seeds in enumerate(gcn_batches(V, T, 8000, rng)):
        if b == num_batches:
            break
        rh = ref.hashtable(2, V)            # direct-addressed table of V buckets (cpu_hashtable2.cc:53-107)
        oh = oracle.hashtable(V + 16)
        rh.populate(seeds)
        oh.fill_unique(seeds)
        cur = seeds
        work = indices.copy()               # CPUSampleKHop2 swaps inside `indices` (cpu_sampling_khop2.cc:29-76)
        for i in (2, 1, 0):                 # cpu_loops.cc:87
            f = fanout[i]
            src, dst = ref.sample_khop2(indptr, work, cur, f)
            # the sampler's contract: seed-major, min(deg, f) DISTINCT neighbours of each seed
            exp_counts = np.minimum(deg[cur], f)
            assert len(src) == int(exp_counts.sum())
            assert np.array_equal(src, np.repeat(cur, exp_counts))
            rh.populate(dst)
            oh.fill_duplicates(dst)
            unique_ref, unique_or = rh.map_nodes(), oh.unique()
            assert np.array_equal(unique_ref, unique_or), "layer %d: ordered unique list differs" % i
            new_src_ref, new_dst_ref = rh.map_edges(src, dst)
            assert np.array_equal(new_src_ref, oh.map(src)) and np.array_equal(new_dst_ref, oh.map(dst))
            assert new_src_ref.max(initial=0) < len(cur) and new_dst_ref.max(initial=0) < len(unique_ref)
            assert np.array_equal(unique_ref[:len(cur)], cur)          # seeds keep their local ids
            cur = unique_ref
        assert np.array_equal(np.sort(work), np.sort(indices))            # khop2 only permutes rows
        # extraction (DoFeatureExtract): rows of the host tables, bit-exact
        assert np.array_equal(ref.extract(feat, cur).view(np.uint32), oracle.extract(feat, cur).view(np.uint32))
        assert np.array_equal(ref.extract(label, seeds), oracle.extract(label, seeds))
        assert np.array_equal(oracle.extract(label, seeds), label[seeds])
