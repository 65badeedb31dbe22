"""Child-process driver for the runtime tests: the engine is a process-wide singleton that can be
configured once, so every scenario runs in its own interpreter.

  python tests/runtime_driver.py <scenario.json> <out.npz>
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))


def collect(sam, key, L, out, tag, with_data):
    import torch
    rec = {}
    for i in range(L):
        rec["row%d" % i] = sam.get_graph_row(key, i).cpu().numpy()
        rec["col%d" % i] = sam.get_graph_col(key, i).cpu().numpy()
        if with_data:
            rec["data%d" % i] = sam.get_graph_data(key, i).cpu().numpy()
        indptr, indices, eids = sam.get_graph_csc(key, i)          # CSC hand-off
        rec["csc_indptr%d" % i] = indptr.cpu().numpy()
        rec["csc_indices%d" % i] = indices.cpu().numpy()
        rec["csc_eids%d" % i] = np.zeros(0, np.int32) if eids is None else eids.cpu().numpy()
        rec["csc_identity%d" % i] = np.array(eids is None)
        rec["nsrc%d" % i] = np.array(sam.get_graph_num_src(key, i))
        rec["ndst%d" % i] = np.array(sam.get_graph_num_dst(key, i))
    rec["feat"] = sam.get_graph_feat(key).cpu().numpy()
    rec["label"] = sam.get_graph_label(key).cpu().numpy()
    rec["input_nodes"] = sam.get_graph_input_nodes(key).cpu().numpy()
    rec["output_nodes"] = sam.get_graph_output_nodes(key).cpu().numpy()
    for k, v in rec.items():
        out["%s/%d/%s" % (tag, key, k)] = v
    torch.cuda.synchronize()


def run_single(sc, out_path):
    import samgraph.torch as sam
    cfg = sc["config"]
    sam.config(cfg)
    sam.init()
    L = cfg.get("num_layer", cfg.get("num_fanout", 0))
    out = {"num_step": np.array(sam.steps_per_epoch()), "num_epoch": np.array(sam.num_epoch()),
           "feat_dim": np.array(sam.feat_dim()), "num_class": np.array(sam.num_class())}
    keys = []
    if sc.get("pipeline"):
        sam.start()
    for epoch in range(sam.num_epoch()):
        for step in range(sam.steps_per_epoch()):
            if not sc.get("pipeline"):
                sam.sample_once()
            key = sam.get_next_batch()
            keys.append(key)
            collect(sam, key, L, out, "b", cfg["_sample_type"] == sam.kRandomWalk)
            out["miss/%d" % key] = np.array(sam.get_log_step_value(epoch, key % sam.steps_per_epoch(), sam.kLogL1MissBytes))
    out["keys"] = np.array(keys, dtype=np.uint64)
    out["sample_time"] = np.array(sam.get_log_epoch_value(0, sam.kLogEpochSampleTime))
    sam.shutdown()
    np.savez(out_path, **out)


def run_arch5(sc, out_path):
    import multiprocessing as mp
    import samgraph.torch as sam
    cfg = sc["config"]
    S, T = cfg["num_sample_worker"], cfg["num_train_worker"]
    sam.config(cfg)
    sam.data_init()
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(S + T, timeout=90)
    L = cfg.get("num_layer", cfg.get("num_fanout", 0))
    num_epoch, num_step = sam.num_epoch(), sam.steps_per_epoch()

    def sampler(wid):
        sam.sample_init(wid, sc["sample_devices"][wid])
        barrier.wait()                                   # samplers ready (presample done)
        barrier.wait()                                   # trainers ready
        for _ in range(num_epoch):
            for _ in range(sam.num_local_step()):
                sam.sample_once()
        barrier.wait()
        sam.shutdown()

    def trainer(wid):
        barrier.wait()
        sam.train_init(wid, sc["train_devices"][wid])
        barrier.wait()
        out = {}
        keys = []
        for epoch in range(num_epoch):
            for step in range(wid, num_step, T):
                sam.sample_once()
                key = sam.get_next_batch()
                keys.append(key)
                collect(sam, key, L, out, "b", cfg["_sample_type"] == sam.kRandomWalk)
        out["keys"] = np.array(keys, dtype=np.uint64)
        np.savez(out_path + ".t%d.npz" % wid, **out)
        barrier.wait()
        sam.shutdown()

    procs = [ctx.Process(target=sampler, args=(i,)) for i in range(S)] + \
            [ctx.Process(target=trainer, args=(i,)) for i in range(T)]
    for p in procs:
        p.start()
    import time
    bad = 0
    deadline = time.time() + 150          # a wedged queue must not eat the GPU budget: kill the stragglers
    for p in procs:
        p.join(max(1.0, deadline - time.time()))
        if p.is_alive():
            p.kill()
            p.join(10)
        bad |= (p.exitcode != 0)
    np.savez(out_path, num_step=np.array(num_step), num_epoch=np.array(num_epoch), bad=np.array(int(bad)))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    sc = json.load(open(sys.argv[1]))
    {"single": run_single, "arch5": run_arch5}[sc["mode"]](sc, sys.argv[2])
