#!/usr/bin/env python
"""End-to-end leg of bench.py: the hot path through the public API (samgraph.torch over the samgraph_*
C-ABI of the C++ engine), one process, dataset loaded from the reference's on-disk format.

  python tools/e2e_runtime.py <dataset_dir> <steps> <warmup> <cache_pct> <device> <seed> [sample_type] [fanout,...]

Per timed step: (sam.sample_once();) key = sam.get_next_batch(); the batch label tensor and the per-layer
edge counts are read back to the host.  Prints one JSON object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def main():
    path, steps, warmup, cache_pct, dev, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), \
        sys.argv[5], int(sys.argv[6])
    sample_type = sys.argv[7] if len(sys.argv) > 7 else "khop2"
    fanout = [int(x) for x in sys.argv[8].split(",")] if len(sys.argv) > 8 else [25, 10]
    import torch
    import samgraph.torch as sam
    meta = dict(l.split() for l in open(os.path.join(path, "meta.txt")) if l.strip())
    spe_guess = (int(meta["NUM_TRAIN_SET"]) + 7999) // 8000
    # enough epochs for warm-up + timed steps + the batches the pipelined sampler keeps in flight
    # (the profiler keeps one record per (epoch, step), so this is sized, not "infinite")
    num_epoch = (steps + warmup + 32) // spe_guess + 2
    cfg = {"dataset_path": path, "_arch": sam.kArch3, "_sample_type": sam.sample_types[sample_type], "batch_size": 8000,
           "num_epoch": num_epoch, "_cache_policy": sam.kCacheByPreSample, "cache_percentage": cache_pct,
           "max_sampling_jobs": 10, "max_copying_jobs": 2, "omp_thread_num": os.cpu_count() or 1,
           "sampler_ctx": dev, "trainer_ctx": dev, "presample_epoch": 1, "seed": seed}
    if sample_type == "random_walk":
        cfg.update(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5, num_layer=3)
        fanout = [5, 5, 5]
    else:
        cfg.update(fanout=fanout, num_fanout=len(fanout), num_layer=len(fanout))
    t0 = time.time()
    sam.config(cfg)
    sam.init()
    init_s = time.time() - t0
    torch.cuda.set_device(torch.device(dev))
    L = len(fanout)
    pipeline = os.environ.get("FGNN_E2E_PIPELINE", "1") != "0"
    if pipeline:
        sam.start()                                            # background sampler + extractor threads (--pipeline)
    for _ in range(warmup):
        if not pipeline:
            sam.sample_once()
        sam.get_next_batch()
    torch.cuda.synchronize()
    edges = 0
    n_in = 0
    d2h = 0
    miss_bytes = 0.0
    spe = sam.steps_per_epoch()
    step_wall = []
    t0 = time.perf_counter()
    t_prev = t0
    for k in range(steps):
        if not pipeline:
            sam.sample_once()
        key = sam.get_next_batch()
        label = sam.get_graph_label(key).cpu()                 # device -> host read of the step's result
        feat = sam.get_graph_feat(key)
        for i in range(L):
            edges += sam.get_graph_num_edge(key, i)
        n_in += feat.shape[0]
        d2h += label.numel() * 8 + L * 3 * 4
        miss_bytes += sam.get_log_step_value(key // spe, key % spe, sam.kLogL1MissBytes)
        t_now = time.perf_counter()
        step_wall.append(t_now - t_prev)
        t_prev = t_now
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"value": edges / dt, "unit": "edges/s", "ms_per_step": dt / steps * 1e3,
           "h2d_bytes_per_step": int(miss_bytes / steps), "d2h_bytes_per_step": d2h // steps,
           "input_nodes_per_step": n_in / steps, "edges_per_step": edges / steps, "init_s": init_s,
           "cache_percentage": cache_pct, "mode": "pipeline (sam.start)" if pipeline else "sample_once",
           "api": "samgraph.torch (start|sample_once / get_next_batch / get_graph_*) over the samgraph_* C-ABI, "
                  "C++ engine arch3",
           "note": "host->device bytes per step = feature rows missing from the HBM cache, read from the pinned host "
                   "feature table inside the gather kernel; device->host = labels + edge counts; host clock around "
                   "the loop"}
    sw = sorted(step_wall)
    out["steps"] = steps
    out["warmup"] = warmup
    out["step_wall_us_p50_p99_max"] = [round(sw[len(sw) // 2] * 1e6, 1), round(sw[min(len(sw) - 1, int(len(sw) * 0.99))] * 1e6, 1),
                                       round(sw[-1] * 1e6, 1)]
    if os.environ.get("FGNN_E2E_DIAG"):
        names = ["kLogL1SampleTime", "kLogL2ShuffleTime", "kLogL2CoreSampleTime", "kLogL1CopyTime",
                 "kLogL2GraphCopyTime", "kLogL2CacheCopyTime", "kLogL2IdCopyTime", "kLogL2ExtractTime"]
        step_ms = sorted(step_wall)
        out["step_wall_us_p50_p99_max"] = [round(step_ms[len(step_ms) // 2] * 1e6, 1),
                                           round(step_ms[int(len(step_ms) * 0.99)] * 1e6, 1), round(step_ms[-1] * 1e6, 1)]
        diag = {}
        for nm in names:
            item = getattr(sam, nm)
            vals = [sam.get_log_step_value((warmup + k) // spe, (warmup + k) % spe, item) for k in range(steps)]
            diag[nm] = round(sum(vals) / len(vals) * 1e6, 1)
            if nm in ("kLogL2IdCopyTime", "kLogL2ExtractTime", "kLogL1CopyTime"):
                sv = sorted(vals)
                diag[nm + "_p50_max"] = [round(sv[len(sv) // 2] * 1e6, 1), round(sv[-1] * 1e6, 1)]
                diag[nm + "_n_over_1ms"] = sum(1 for v in vals if v > 1e-3)
        out["diag_us"] = diag
    print("E2E_JSON " + json.dumps(out))
    sam.shutdown()


if __name__ == "__main__":
    main()
