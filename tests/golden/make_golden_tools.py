#!/usr/bin/env python
"""Generate tests/golden/ref_tools_golden.npz from the REFERENCE's own offline dataset tools
(utility/data-process/toolkit/{weight,cache}/*.cc, compiled in place by `make -C oracle reftools` into
oracle/_ref/tools/).  They are run on a small dataset written in the reference's on-disk format; the fixture holds
their outputs so that the oracle's restatements (and, through the oracle, the GPU builders) are pinned against the
real reference code on boxes without /root/reference.

  python tests/golden/make_golden_tools.py

Weight policies: kDefault draws from std::random_device (not reproducible), so the two deterministic policies of
the same tools are used — kInverseSrcDegreeRand (w = 1 / out_degree[src]) and kSrcSuffix (w = 100 if
out_degree[src] < 10 else 1); create_alias_table.cc:75-92, create_prob_prefix_table.cc:73-90.
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
SPEC = (3000, 24000, 4, 5, 400)      # V, E, feat_dim, classes, train nodes
SEED = 21


def run(tool, root, *extra):
    # one thread: the tools are deterministic per row either way, this only keeps the run quiet and small
    cmd = [os.path.join(TOOLS, tool), "-p", root, "-g", "products", "-t", "1"] + list(extra)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftools"])
    ds = make_dataset_numpy(SPEC, seed=SEED)
    root = tempfile.mkdtemp(prefix="fgnn_reftools_")
    folder = os.path.join(root, "products")
    write_dataset(folder, ds)
    out = {"spec": np.array(SPEC), "seed": np.array(SEED),
           # enough to detect a drift of the synthetic generator itself
           "indptr_sum": np.array(int(ds["indptr"].astype(np.int64).sum())),
           "indices_sum": np.array(int(ds["indices"].astype(np.int64).sum())),
           "train_sum": np.array(int(ds["train_set"].astype(np.int64).sum()))}
    for policy in ("kInverseSrcDegreeRand", "kSrcSuffix"):
        run("create_alias_table", root, "-P", policy)
        run("create_prob_prefix_table", root, "-P", policy)
        out["prob_" + policy] = np.fromfile(os.path.join(folder, "prob_table.bin"), np.float32)
        out["alias_" + policy] = np.fromfile(os.path.join(folder, "alias_table.bin"), np.uint32)
        out["prefix_" + policy] = np.fromfile(os.path.join(folder, "prob_prefix_table.bin"), np.float32)
    run("cache_by_degree", root)
    run("cache_by_heuristic", root)
    out["cache_by_degree"] = np.fromfile(os.path.join(folder, "cache_by_degree.bin"), np.uint32)
    out["cache_by_heuristic"] = np.fromfile(os.path.join(folder, "cache_by_heuristic.bin"), np.uint32)
    shutil.rmtree(root)
    path = os.path.join(HERE, "ref_tools_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape, v.dtype) for k, v in out.items() if v.ndim})


if __name__ == "__main__":
    main()
