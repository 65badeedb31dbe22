#!/usr/bin/env python
"""A/B sweep of the end-to-end leg (tools/e2e_runtime.py: samgraph.torch over the C++ engine) at the bench.py
shape: the dataset is generated and written once, then the engine runs once per environment variant.

  python tools/e2e_sweep.py [--cache-pct 0.25] [--steps 302] VAR=val[,VAR2=val2] ...
e.g. python tools/e2e_sweep.py base FGNN_SUPER_BATCH=1 FGNN_SUPER_BATCH=8 FGNN_EXTRACT_PRIORITY=0 FGNN_BULK_STAGES=3
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="papers100M")
    ap.add_argument("--cache-pct", type=float, default=0.25)
    ap.add_argument("--steps", type=int, default=302)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--empty-feat", type=int, default=22)
    ap.add_argument("variants", nargs="*", default=["base"])
    a = ap.parse_args()
    import torch
    import bench
    a.sample_type, a.fanout = "khop2", "25,10"
    wl = bench.build_workload(a, "cuda:0")
    path = bench.write_dataset_shm(a, wl, 0, 1)
    del wl
    torch.cuda.empty_cache()
    for var in a.variants:
        env = dict(os.environ, SAMGRAPH_EMPTY_FEAT=str(a.empty_feat), SAMGRAPH_LOG_LEVEL="error", FGNN_E2E_DIAG="1")
        if var != "base":
            for kv in var.split(","):
                k, v = kv.split("=", 1)
                env[k] = v
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "e2e_runtime.py"), path, str(a.steps),
                            str(a.warmup), str(a.cache_pct), "cuda:0", str(0x5EED0000), "khop2", "25,10"],
                           capture_output=True, text=True, env=env, timeout=900)
        res = None
        for line in r.stdout.splitlines():
            if line.startswith("E2E_JSON "):
                res = json.loads(line[len("E2E_JSON "):])
        if res is None:
            print("%-40s FAILED %s" % (var, r.stderr[-300:]))
            continue
        gbps = res["h2d_bytes_per_step"] / (res["ms_per_step"] * 1e-3) / 1e9
        print("%-40s ms/step %.4f  %.3f G edges/s  p50/p99/max us %s  host-link %.1f GB/s  diag %s" % (
            var, res["ms_per_step"], res["value"] / 1e9, res["step_wall_us_p50_p99_max"], gbps,
            {k: v for k, v in res.get("diag_us", {}).items() if k in ("kLogL2ExtractTime", "kLogL2IdCopyTime")}), flush=True)
    import shutil
    shutil.rmtree(path, ignore_errors=True)


if __name__ == "__main__":
    main()
