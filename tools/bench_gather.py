#!/usr/bin/env python
"""Micro-benchmark + self-check of fgnn_k_gather_cached at the bench.py shape.

  python tools/bench_gather.py [--rows N_in] [--dim D] [--cache-rows C] [--miss-frac f] [--iters K] [--sweep]

One process = one kernel configuration (the FGNN_GATHER_* switches are read once at
first launch); --sweep forks one child per configuration and prints a table.
Timing: CUDA events around K back-to-back launches on K different id lists
(each list gathers N_in*D*4 B > L2, cache table much larger than L2).
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def one(args, cfgs=({},)):
    import torch
    from fgnn_b200 import kernels as K
    K.load()
    dev = "cuda"
    D4 = args.dim * 4
    V = args.num_nodes
    C = args.cache_rows
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    cache = torch.randint(-2**31, 2**31 - 1, (C, args.dim), dtype=torch.int32, device=dev, generator=g)
    # node -> slot: first nodes of a random permutation are cached
    table = torch.full((V,), -1, dtype=torch.int32, device=dev)
    perm = torch.randperm(V, device=dev, generator=g)
    cached_nodes = perm[:C]
    table[cached_nodes] = torch.arange(C, dtype=torch.int32, device=dev)
    host_rows = 1 << 20
    host = torch.randint(-2**31, 2**31 - 1, (host_rows, args.dim), dtype=torch.int32).pin_memory()
    n = args.rows
    n_miss = int(n * args.miss_frac)
    lists = []
    for k in range(args.iters):
        hit_ids = cached_nodes[torch.randint(0, C, (n - n_miss,), device=dev, generator=g)]
        if n_miss:
            miss_ids = perm[C + torch.randint(0, V - C, (n_miss,), device=dev, generator=g)]
            ids = torch.cat([hit_ids, miss_ids])[torch.randperm(n, device=dev, generator=g)]
        else:
            ids = hit_ids
        lists.append(ids.to(torch.int32).contiguous())
    out = torch.empty((n, args.dim), dtype=torch.int32, device=dev)
    d_n = torch.tensor([n], dtype=torch.int32, device=dev)
    ptrs = torch.tensor([cache.data_ptr()], dtype=torch.int64, device=dev)
    stats = torch.zeros(2, dtype=torch.int64, device=dev)

    def run(ids):
        K.gather_cached(out, ids, n, d_n, table, ptrs, 1, host, D4, stats, host_rows - 1)

    rc = 0
    for cfg in cfgs:
        for k in [k for k in os.environ if k.startswith("FGNN_") and k != "FGNN_TUNING_DYNAMIC" and cfgs != ({},)]:
            del os.environ[k]
        os.environ.update(cfg)
        try:
            out.zero_()
            stats.zero_()
            run(lists[0])
            torch.cuda.synchronize()
            ids = lists[0].long()
            slot = table[ids].long()
            exp = torch.where((slot >= 0).unsqueeze(1), cache[slot.clamp(min=0)],
                              host.to(dev)[ids & (host_rows - 1)])
            ok = bool(torch.equal(out, exp))
            st = stats.tolist()
            ok = ok and st[0] == n - n_miss and st[1] == n_miss
            del exp
            for _ in range(3):
                run(lists[1 % args.iters])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for rep in range(3):
                e0.record()
                for ids in lists:
                    run(ids)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / args.iters)
            alg = n * (4 + 2 * D4)
            print("GATHER_JSON " + json.dumps({"ok": ok, "us": round(best * 1e3, 2), "GBps": round(alg / best / 1e6, 1),
                                               "cfg": cfg}))
            rc |= 0 if ok else 1
        except Exception as e:  # a configuration that cannot launch (e.g. shared memory) is reported, not fatal
            print("GATHER_JSON " + json.dumps({"ok": False, "error": str(e)[:200], "cfg": cfg}))
            rc |= 2
        sys.stdout.flush()
    return rc


def sweep(args):
    """All configurations in ONE process (FGNN_TUNING_DYNAMIC=1 makes the library re-read the switches)."""
    os.environ["FGNN_TUNING_DYNAMIC"] = "1"
    cfgs = [{"FGNN_GATHER_IMPL": "flat"}, {"FGNN_GATHER_IMPL": "group"}]
    if args.miss_frac > 0:
        cfgs += [{"FGNN_GATHER_IMPL": "bulk", "FGNN_BULK_MISS_LDG": "1"},
                 {"FGNN_GATHER_IMPL": "bulk", "FGNN_BULK_MISS_LDG": "0"}]
    else:
        for nw, s_, sb in ((8, 6, 2048), (8, 6, 4096), (8, 8, 2048), (8, 8, 1024), (8, 12, 2048), (8, 12, 1024),
                           (4, 8, 2048), (4, 8, 4096), (4, 12, 2048), (4, 12, 4096), (4, 16, 2048), (4, 16, 1024),
                           (16, 6, 2048), (16, 4, 2048), (16, 8, 1024), (16, 3, 4096), (16, 4, 1024)):
            cfgs.append({"FGNN_GATHER_IMPL": "bulk", "FGNN_BULK_WARPS": str(nw), "FGNN_BULK_STAGES": str(s_),
                         "FGNN_BULK_STAGE_BYTES": str(sb)})
    one(args, cfgs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=550000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--num-nodes", type=int, default=111059956)
    ap.add_argument("--cache-rows", type=int, default=8000000)
    ap.add_argument("--miss-frac", type=float, default=0.0)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--sweep", action="store_true")
    a = ap.parse_args()
    if a.sweep:
        sweep(a)
    else:
        sys.exit(one(a))
