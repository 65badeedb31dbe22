#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own CPU code (oracle/_ref, compiled from
/root/reference by `make -C oracle ref`).  Run in the build container; the fixtures are committed so the
GPU box (which has no /root/reference) can check the oracle and the CUDA path against reference outputs.

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fgnn-artifacts_b200")]

from fgnn_b200.synth import make_graph_numpy  # noqa: E402
from oracle.oracle import RefCPU, build_ref  # noqa: E402


def main():
    build_ref()
    ref = RefCPU()
    ref.set_threads(1)     # a 1-thread CPUHashTable2 is deterministic (first occurrence wins the CAS)
    rng = np.random.default_rng(2024)
    indptr, indices = make_graph_numpy(4000, 60000, seed=5)
    V = len(indptr) - 1

    # 1. unique / remap: CPUHashTable0 and CPUHashTable2 (cpu_hashtable0.cc, cpu_hashtable2.cc)
    seeds = rng.permutation(V)[:300].astype(np.uint32)
    rounds = [rng.integers(0, V, size=n).astype(np.uint32) for n in (2000, 7000, 15000)]
    out = {"ht_seeds": seeds}
    for kind in (0, 2):
        ht = ref.hashtable(kind, V if kind == 2 else 40000)
        ht.populate(seeds)
        for r, ids in enumerate(rounds):
            ht.populate(ids)
            out["ht%d_unique_%d" % (kind, r)] = ht.map_nodes()
            ns, nd = ht.map_edges(ids, ids[::-1].copy())
            out["ht%d_map_src_%d" % (kind, r)] = ns
            out["ht%d_map_dst_%d" % (kind, r)] = nd
    for r, ids in enumerate(rounds):
        out["ht_round_%d" % r] = ids

    # 2. CPUExtract (cpu_extraction.cc) for the dtypes the engine moves
    feat = rng.standard_normal((V, 24)).astype(np.float32)
    label = rng.integers(0, 47, size=V).astype(np.int64)
    idx = rng.integers(0, V, size=5000).astype(np.uint32)
    out.update(ex_feat=feat, ex_label=label, ex_index=idx, ex_feat_out=ref.extract(feat, idx),
               ex_label_out=ref.extract(label, idx))

    # 3. CPUSampleKHop0/2 (cpu_sampling_khop0.cc, cpu_sampling_khop2.cc): the deterministic part — rows with
    #    degree <= fanout are copied whole, out_src and the per-row counts are fixed by the degrees alone.
    inp = rng.permutation(V)[:1500].astype(np.uint32)
    for fanout in (5, 15):
        s0, d0 = ref.sample_khop0(indptr, indices, inp, fanout)
        s2, d2 = ref.sample_khop2(indptr, indices.copy(), inp, fanout)
        out["khop_src_f%d" % fanout] = s0
        assert np.array_equal(s0, s2)
        deg = np.diff(indptr.astype(np.int64))[inp]
        small = np.repeat(deg <= fanout, np.minimum(deg, fanout))
        out["khop_small_mask_f%d" % fanout] = small
        out["khop_small_dst_f%d" % fanout] = d0[small]
        assert np.array_equal(d0[small], d2[small])
    out.update(g_indptr=indptr, g_indices=indices, khop_input=inp)

    # 4. sampling distribution of the reference (seeded by its own thread-local mt19937): per-neighbour pick
    #    frequencies over many draws of one row, for the chi-square test
    hub = int(np.argmax(np.diff(indptr.astype(np.int64))))
    reps = 4000
    one = np.full(reps, hub, dtype=np.uint32)
    for name, fn in (("khop0", lambda: ref.sample_khop0(indptr, indices, one, 10)),
                     ("khop2", lambda: ref.sample_khop2(indptr, indices.copy(), one, 10))):
        s, d = fn()
        row = indices[indptr[hub]:indptr[hub + 1]]
        # khop2 permutes the row in place, but the multiset of the row is invariant
        ids, cnt = np.unique(d, return_counts=True)
        out["dist_%s_ids" % name] = ids
        out["dist_%s_cnt" % name] = cnt
    out["dist_hub"] = np.array(hub)
    out["dist_reps"] = np.array(reps)

    # 5. PredictNumNodes (common.cc:330-339)
    out["predict_8000_25_10"] = np.array(ref.predict_num_nodes(8000, [25, 10]))
    out["predict_8000_5_10_15"] = np.array(ref.predict_num_nodes(8000, [5, 10, 15]))

    np.savez_compressed(os.path.join(HERE, "ref_cpu_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_cpu_golden.npz"), "with", len(out), "arrays")


if __name__ == "__main__":
    main()
