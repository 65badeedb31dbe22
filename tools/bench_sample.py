#!/usr/bin/env python
"""Micro-benchmark of the sampling chain (fgnn_k_sample_batch) at the bench.py shape, alone and overlapped
with the feature gather, for the kernel-fusion variants (FGNN_BATCH_FUSE bit 0: insert while sampling,
bit 1: remap folded into the compaction).

  python tools/bench_sample.py [--steps 150]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=150)
    ap.add_argument("--workload", default="papers100M")
    ap.add_argument("--empty-feat", type=int, default=22)
    ap.add_argument("--cache-pct", type=float, default=0.3)
    ap.add_argument("--sweep", default="fuse", choices=["fuse", "overlap", "profile", "diag", "timeline"],
                    help="fuse: kernel-fusion variants; overlap: how the sampling slots and the gather share the GPU\n                    (FGNN_GRID_DIV, gather implementation / CTA shape)")
    a = ap.parse_args()
    import torch
    import bench
    from fgnn_b200 import kernels as K
    from fgnn_b200.pipeline import HotPath
    dev = "cuda:0"
    torch.cuda.set_device(0)
    K.load()
    wl = bench.build_workload(a, dev)
    V, D = wl["V"], wl["D"]
    BATCH, FANOUTS = bench.BATCH, bench.FANOUTS
    spe = (wl["T"] + BATCH - 1) // BATCH
    perm = wl["train"]
    SL = 6
    hp = HotPath(wl["indptr"], wl["indices"], V, FANOUTS, BATCH, "khop2", seed=1, device=dev, num_slots=SL,
                 ht_capacity=(1 << int(os.environ["FGNN_DIAG_HT_LOG2"])) if os.environ.get("FGNN_DIAG_HT_LOG2") else None)
    rank = torch.randperm(V, device=dev).to(torch.int32)
    hp.build_cache(rank, a.cache_pct, wl["host_feat"], D * 4, wl["feat_mask"])
    # every row "hit": table over the cached fraction only -> make all nodes map into the cache (mod)
    hp.cache_table = (torch.arange(V, device=dev, dtype=torch.int64) % hp.num_cached).to(torch.int32)
    hp.set_labels(wl["label"])
    streams = [torch.cuda.Stream(device=dev) for _ in range(SL)]
    xs = torch.cuda.Stream(device=dev, priority=-1)

    def seeds_of(k):
        s = k % (spe - 1)
        return perm[s * BATCH:(s + 1) * BATCH], BATCH

    host_us = [0.0]
    copy_src = torch.empty((560000, D * 4), dtype=torch.uint8, device=dev)

    def run(slots, with_gather, steps, copy_instead=False, gather_only=False):
        sampled = [torch.cuda.Event() for _ in range(slots)]
        gathered = [torch.cuda.Event() for _ in range(slots)]
        main = torch.cuda.current_stream()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for st in streams + [xs]:
            st.wait_stream(main)
        e0.record()
        for st in streams + [xs]:
            st.wait_stream(main)
        th0 = time.perf_counter()
        for k in range(steps):
            sd, n = seeds_of(k)
            sl = k % slots
            if not gather_only:
                with torch.cuda.stream(streams[sl]):
                    if with_gather:
                        streams[sl].wait_event(gathered[sl])
                    hp.sample(sd, n, 7000 + k, slot=sl)
                    sampled[sl].record()
            if with_gather:
                with torch.cuda.stream(xs):
                    if not gather_only:
                        xs.wait_event(sampled[sl])
                    if copy_instead:
                        hp.feat_out[:560000].copy_(copy_src)   # pure streaming copy of the gather's size
                    else:
                        hp.gather(sl)
                    hp.gather_labels(sd, n)
                    gathered[sl].record()
        host_us[0] = (time.perf_counter() - th0) / steps * 1e6
        for st in streams + [xs]:
            main.wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps * 1e3

    if a.sweep == "profile":
        # for `ncu --profile-from-start off`: one slot, sampling chain only, environment as given
        run(1, False, 10)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        us = run(1, False, a.steps)
        torch.cuda.profiler.stop()
        print("PROFILE_JSON " + json.dumps({"sample_only_us_slots1": round(us, 1)}))
        return
    if a.sweep == "timeline":
        # who runs when in the overlapped loop: CUDA events after every launch (fgnn_k_trace_*), 4 slots + gather
        out = {}
        for tag, slots, with_gather in (("sample_only_1slot", 1, False), ("sample_only_4slots", 4, False),
                                        ("with_gather_4slots", 4, True)):
            run(slots, with_gather, 12)
            torch.cuda.synchronize()
            K.trace_enable(True, 4096)
            run(slots, with_gather, 12)
            recs = K.trace_dump(4096)
            # per stream: duration of every launch = its end mark minus the previous mark on that stream
            by_stream, rows = {}, []
            for name, st, t in recs:
                prev = by_stream.get(st)
                by_stream[st] = t
                if name in ("batch_begin", "gather_begin") or prev is None:
                    continue
                rows.append((round(prev * 1e3, 1), round(t * 1e3, 1), name, st))
            rows.sort()
            sids = {s: i for i, s in enumerate(sorted({r[3] for r in rows}))}
            agg = {}
            for b, e, name, st in rows:
                agg.setdefault(name, []).append(e - b)
            out[tag] = {"per_kernel_us_mean": {k: round(sum(v) / len(v), 1) for k, v in sorted(agg.items())},
                        "first_200_intervals_us": [[b, e, name, sids[st]] for b, e, name, st in rows[:200]]}
        print("TIMELINE_JSON " + json.dumps(out))
        return
    os.environ["FGNN_TUNING_DYNAMIC"] = "1"
    if a.sweep == "diag":
        # what bounds the overlapped loop?  host enqueue rate, the gather's SM footprint, or shared DRAM time
        res = {"ht_log2": os.environ.get("FGNN_DIAG_HT_LOG2"), "ht_MB_per_slot": K.ht_bytes(hp.cap) / 1e6}
        for var in ({}, {"FGNN_BULK_WARPS": "8", "FGNN_BULK_STAGES": "4"}, {"FGNN_BULK_WARPS": "16", "FGNN_BULK_STAGES": "3"},
                    {"FGNN_BULK_WARPS": "4", "FGNN_BULK_STAGES": "8"}):
            os.environ.update({"FGNN_BULK_WARPS": "16", "FGNN_BULK_STAGES": "6"})
            os.environ.update(var)
            tag = "w%ss%s" % (os.environ["FGNN_BULK_WARPS"], os.environ["FGNN_BULK_STAGES"])
            run(4, True, 10)
            res["with_gather_us_slots4_" + tag] = round(run(4, True, a.steps), 1)
            res["host_enqueue_us_" + tag] = round(host_us[0], 1)
            run(1, True, 10, gather_only=True)
            res["gather_only_us_" + tag] = round(run(1, True, a.steps, gather_only=True), 1)
        os.environ.update({"FGNN_BULK_WARPS": "16", "FGNN_BULK_STAGES": "6"})
        for slots in (1, 2, 3, 4, 6):
            run(slots, False, 10)
            res["sample_only_us_slots%d" % slots] = round(run(slots, False, a.steps), 1)
            res["sample_only_host_us_slots%d" % slots] = round(host_us[0], 1)
        run(4, True, 10, copy_instead=True)
        res["with_copy_us_slots4"] = round(run(4, True, a.steps, copy_instead=True), 1)
        run(1, True, 10, copy_instead=True, gather_only=True)
        res["copy_only_us"] = round(run(1, True, a.steps, copy_instead=True, gather_only=True), 1)
        res["n_in"] = int(hp.slots[0].num_items.item())
        print("DIAG_JSON " + json.dumps(res))
        return
    if a.sweep == "overlap":
        base = {"FGNN_GRID_DIV": "1", "FGNN_GATHER_IMPL": "bulk", "FGNN_BULK_WARPS": "8", "FGNN_BULK_STAGES": "8",
                "FGNN_GATHER_CTAS_PER_SM": "0", "FGNN_BATCH_FUSE": "2"}
        dyn16 = {"FGNN_GATHER_IMPL": "dyn", "FGNN_BULK_WARPS": "16", "FGNN_BULK_STAGES": "6"}
        w16 = {"FGNN_BULK_WARPS": "16", "FGNN_BULK_STAGES": "6"}
        variants = [{}, {"FGNN_BATCH_FUSE": "6"}, dict(w16), dict(w16, FGNN_BATCH_FUSE="6"),
                    dict(dyn16, FGNN_BATCH_FUSE="6"), dict(w16, FGNN_BATCH_FUSE="6", FGNN_GRID_DIV="2")]
        for var in variants:
            env = dict(base, **var)
            os.environ.update(env)
            res = dict(var)
            run(1, False, 10)
            res["sample_only_us_slots1"] = round(run(1, False, a.steps), 1)
            res["sample_only_us_slots3"] = round(run(3, False, a.steps), 1)
            for slots in (2, 3, 4, 6):
                run(slots, True, 10)
                res["with_gather_us_slots%d" % slots] = round(run(slots, True, a.steps), 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for k in range(5):
                hp.gather(0)
            e0.record()
            for k in range(50):
                hp.gather(0)
            e1.record()
            torch.cuda.synchronize()
            res["gather_alone_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
            print("OVERLAP_JSON " + json.dumps(res))
            sys.stdout.flush()
        return
    for fuse, hint in ((0, 0), (0, 1), (2, 0), (2, 1), (3, 1)):
        os.environ["FGNN_BATCH_FUSE"] = str(fuse)
        os.environ["FGNN_GATHER_L2HINT"] = str(hint)
        res = {"fuse": fuse, "l2_evict_first": hint}
        for slots in (1, 3):
            run(slots, False, 10)
            res["sample_only_us_slots%d" % slots] = round(run(slots, False, a.steps), 1)
        for slots in (1, 2, 3, 4):
            run(slots, True, 10)
            res["with_gather_us_slots%d" % slots] = round(run(slots, True, a.steps), 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(50):
            hp.gather(0)
        e1.record()
        torch.cuda.synchronize()
        res["gather_alone_us"] = round(e0.elapsed_time(e1) / 50 * 1e3, 1)
        print("SAMPLE_JSON " + json.dumps(res))
        sys.stdout.flush()
    os.environ.pop("FGNN_BATCH_FUSE")
    # gather alone for reference
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(50):
        hp.gather(0)
    e1.record()
    torch.cuda.synchronize()
    print("SAMPLE_JSON " + json.dumps({"gather_alone_us": round(e0.elapsed_time(e1) / 50 * 1e3, 1),
                                       "n_in": int(hp.slots[0].num_items.item())}))


if __name__ == "__main__":
    main()
