#!/usr/bin/env python
"""Generate tests/golden/ref_cache_policy_golden.npz from the REFERENCE's own offline cache-ranking tools
toolkit/cache/cache_by_degree_hop.cc and cache_by_fake_optimal.cc (compiled in place by `make -C oracle reftools`).
Same procedure as make_golden_tools.py: a small dataset in the reference's on-disk format, the tools' output files
stored as the fixture.  cache_by_fake_optimal is run with -t 1 and -t 4: its floating-point products are taken in an
order that depends on the thread count (vertices bucketed by id % threads), which the restatements reproduce.

  python tests/golden/make_golden_policies.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fgnn-artifacts_b200")]

from fgnn_b200.synth import make_dataset_numpy, write_dataset  # noqa: E402

TOOLS = os.path.join(ROOT, "oracle", "_ref", "tools")
# cache_by_degree_hop.cc:90 leaves new_indptr[] of the untouched vertices UNINITIALISED (`new uint32_t[n]`): the tool
# only works when that allocation comes back zeroed, i.e. when it is large enough to be mmap()ed by malloc — true for
# every real dataset, not for a 3000-vertex toy (it read garbage, took seconds and crashed with -t 4).  Hence V >= 40k.
CASES = {"a": ((40000, 300000, 4, 5, 60), 21),   # V, E, feat_dim, classes, train nodes ; seed
         "b": ((40000, 900000, 4, 5, 25), 22),   # denser: hubs inside the two-hop ball, many multi-edges
         "c": ((40000, 2000000, 4, 5, 300), 23)}  # two-hop balls overlap: expectations accumulate over many seeds


def run(tool, root, threads):
    subprocess.run([os.path.join(TOOLS, tool), "-p", root, "-g", "products", "-t", str(threads)], check=True,
                   stdout=subprocess.DEVNULL)


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftools"])
    out = {}
    for name, (spec, seed) in CASES.items():
        ds = make_dataset_numpy(spec, seed=seed)
        root = tempfile.mkdtemp(prefix="fgnn_refpolicy_")
        folder = os.path.join(root, "products")
        write_dataset(folder, ds)
        out[name + "_spec"] = np.array(spec)
        out[name + "_seed"] = np.array(seed)
        out[name + "_indices_sum"] = np.array(int(ds["indices"].astype(np.int64).sum()))
        out[name + "_train_sum"] = np.array(int(ds["train_set"].astype(np.int64).sum()))
        for t in (1, 4):
            run("cache_by_degree_hop", root, t)
            out["%s_degree_hop_t%d" % (name, t)] = np.fromfile(os.path.join(folder, "cache_by_degree_hop.bin"), np.uint32)
            run("cache_by_fake_optimal", root, t)
            out["%s_fake_optimal_t%d" % (name, t)] = np.fromfile(os.path.join(folder, "cache_by_fake_optimal.bin"), np.uint32)
        shutil.rmtree(root)
    path = os.path.join(HERE, "ref_cache_policy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if v.ndim})


if __name__ == "__main__":
    main()
