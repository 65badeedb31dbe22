#!/usr/bin/env python
"""Write a synthetic dataset in the reference's on-disk format (meta.txt + indptr/indices/feat/label/*_set .bin,
datagen/README.md, engine.cc:73-264) — the power-law generator bench.py and the tests use.

  python tools/make_dataset.py ci-1m /dev/shm/ci-1m            # a named shape of fgnn_b200.synth.SHAPES
  python tools/make_dataset.py 200000,3000000,64,16,20000 /tmp/ds --weights   # V,E,feat_dim,classes,train nodes

--weights also writes prob_table.bin / alias_table.bin / prob_prefix_table.bin (kDefault weights 1..10) using the
GPU builders (fgnn_k_build_alias_table / fgnn_k_build_prefix_table); it needs a CUDA device.  cache_by_*.bin files
are not needed: the engine ranks the vertices itself for pre_sample / degree / heuristic / random (on the GPU) and for
degree_hop / fake_optimal (on the host, at data_init).  --cache-policy-files writes cache_by_degree_hop.bin and
cache_by_fake_optimal.bin up front with the same host builders (include/fgnn_dataset_tools.h; the equivalents of the
reference's toolkit/cache/cache_by_degree_hop.cc and cache_by_fake_optimal.cc); no GPU needed.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", help="name in fgnn_b200.synth.SHAPES or V,E,feat_dim,num_class,num_train")
    ap.add_argument("out_dir")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--no-feat", action="store_true", help="omit feat.bin (run with SAMGRAPH_EMPTY_FEAT=k)")
    ap.add_argument("--weights", action="store_true")
    ap.add_argument("--cache-policy-files", action="store_true")
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import numpy as np
    from fgnn_b200.synth import SEED, SHAPES, make_dataset_numpy, write_dataset
    shape = a.shape if a.shape in SHAPES else tuple(int(x) for x in a.shape.split(","))
    ds = make_dataset_numpy(shape, seed=SEED if a.seed is None else a.seed, with_feat=not a.no_feat)
    write_dataset(a.out_dir, ds)
    if a.weights:
        import torch
        from fgnn_b200 import kernels as K
        K.load()
        V, E = ds["num_node"], ds["num_edge"]
        indptr = torch.from_numpy(ds["indptr"].view(np.int32)).cuda()
        indices = torch.from_numpy(ds["indices"].view(np.int32)).cuda()
        w = torch.from_numpy(ds["edge_weight"]).cuda()
        prob = torch.empty(E, dtype=torch.float32, device="cuda")
        alias = torch.empty(E, dtype=torch.int32, device="cuda")
        prefix = torch.empty(E, dtype=torch.float32, device="cuda")
        K.build_alias_table(indptr, indices, V, E, w, prob, alias)
        K.build_prefix_table(indptr, V, w, prefix)
        torch.cuda.synchronize()
        prob.cpu().numpy().tofile(os.path.join(a.out_dir, "prob_table.bin"))
        alias.cpu().numpy().tofile(os.path.join(a.out_dir, "alias_table.bin"))
        prefix.cpu().numpy().tofile(os.path.join(a.out_dir, "prob_prefix_table.bin"))
    if a.cache_policy_files:
        import ctypes
        lib = ctypes.CDLL(os.path.join(ROOT, "fgnn-artifacts_b200", "samgraph", "torch", "c_lib.so"))
        P, Z, I = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        lib.fgnn_rt_rank_degree_hop.argtypes = [P, P, Z, P, Z, I, I, P]
        lib.fgnn_rt_rank_fake_optimal.argtypes = [P, P, Z, P, Z, I, I, I, I, P]
        ip, ix, tr = (np.ascontiguousarray(ds[k], np.uint32) for k in ("indptr", "indices", "train_set"))
        rank = np.empty(ds["num_node"], np.uint32)
        assert lib.fgnn_rt_rank_degree_hop(ip.ctypes.data, ix.ctypes.data, len(rank), tr.ctypes.data, len(tr), 2,
                                           a.threads, rank.ctypes.data) == 0
        rank.tofile(os.path.join(a.out_dir, "cache_by_degree_hop.bin"))
        assert lib.fgnn_rt_rank_fake_optimal(ip.ctypes.data, ix.ctypes.data, len(rank), tr.ctypes.data, len(tr), 25, 10,
                                             48, a.threads, rank.ctypes.data) == 0
        rank.tofile(os.path.join(a.out_dir, "cache_by_fake_optimal.bin"))
    print("wrote %s: %d nodes, %d edges, feat_dim %d, %d classes, %d train nodes" %
          (a.out_dir, ds["num_node"], ds["num_edge"], ds["feat_dim"], ds["num_class"], len(ds["train_set"])))


if __name__ == "__main__":
    main()
