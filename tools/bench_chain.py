#!/usr/bin/env python
"""Sampling chain alone (no gather) at the bench.py shape: microseconds per mini-batch for super-batches of
K = 1, 2, 4, 8 mini-batches (fgnn_k_sample_batch_multi), warm (the loop runs back to back, tables and CSR hot
in L2 as far as they fit) with a per-launch breakdown from the event trace.  Knobs come from the environment
(FGNN_BATCH_FUSE, FGNN_HT_VERSIONED, ...), so A/B runs are two invocations.

  python tools/bench_chain.py [--workload papers100M] [--fanout 25,10] [--reps 30] [--ks 1,2,4,8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="papers100M")
    ap.add_argument("--fanout", default="25,10")
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--ks", default="1,2,4,8")
    ap.add_argument("--empty-feat", type=int, default=20)
    ap.add_argument("--sample-type", default="khop2")
    a = ap.parse_args()
    import torch
    import bench
    from fgnn_b200 import kernels as K
    from fgnn_b200.pipeline import HotPath
    torch.cuda.set_device(0)
    K.load()
    a.cache_pct = 0.0
    wl = bench.build_workload(a, "cuda:0")
    fanouts = [int(x) for x in a.fanout.split(",")]
    B = bench.BATCH
    ks = [int(x) for x in a.ks.split(",")]
    hp = HotPath(wl["indptr"], wl["indices"], wl["V"], fanouts, B, a.sample_type, seed=1, device="cuda:0",
                 num_slots=max(ks), rw=bench.RW if a.sample_type == "random_walk" else None,
                 ht_capacity=(1 << int(os.environ["FGNN_DIAG_HT_LOG2"])) if os.environ.get("FGNN_DIAG_HT_LOG2") else None)
    spe = wl["T"] // B
    perm = wl["train"]
    out = {"workload": a.workload, "fanout": fanouts, "env": {k: v for k, v in os.environ.items() if k.startswith("FGNN_")}}
    step = 0
    for Kk in ks:
        def group(g):
            nonlocal step
            batches = []
            for j in range(Kk):
                s = step % spe
                step += 1
                batches.append((perm[s * B:(s + 1) * B], B, 1000 + step, j))
            if Kk == 1:
                hp.sample(*batches[0][:3], slot=0)
            else:
                hp.sample_multi(batches)
        for g in range(3):
            group(g)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for g in range(a.reps):
            group(g)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (a.reps * Kk)
        # per-launch breakdown of a few groups
        K.trace_enable(True, 4096)
        for g in range(4):
            group(g)
        torch.cuda.synchronize()
        rec = K.trace_dump(4096)
        K.trace_enable(False, 0)
        per = {}
        prev = None
        for name, stream, ms in rec:
            if prev is not None and name != "batch_begin":
                per.setdefault(name, []).append((ms - prev) * 1e3)
            prev = ms
        edges = sum(int(hp.slots[j].counts[:, 1].sum().item()) for j in range(Kk)) / Kk
        out["K=%d" % Kk] = {"us_per_batch": round(us, 2), "edges_per_batch": edges,
                            "launch_us_per_group": {k: round(sum(v) / len(v), 1) for k, v in per.items()}}
        print("K=%d  %.1f us per mini-batch   %s" % (Kk, us, out["K=%d" % Kk]["launch_us_per_group"]), flush=True)
    print("CHAIN_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
