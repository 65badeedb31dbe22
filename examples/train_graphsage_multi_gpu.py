#!/usr/bin/env python
"""GraphSAGE / GCN-style training in FGNN's FACTORED mode (dedicated sampler GPUs + dedicated trainer GPUs) on the
samgraph B200 runtime, without DGL: the process structure, barriers and sam.* calls of the reference's
example/samgraph/multi_gpu/train_graphsage.py:95-215 (run_sample) and :217-437 (run_train) — config + data_init in
the parent, fork, sample_init / train_init in the children, per-epoch barriers, extract_start in pipeline mode —
with the blocks taken in CSC form and the mean aggregation written as a sparse-CSR SpMM (examples/
train_graphsage_csc.py).  Placement follows common_config.py:182-185: trainers on cuda:0..T-1 ("trainer gpu id
should start from 0"), samplers on cuda:T..T+S-1 (--single-gpu: everything on cuda:0).

  PYTHONPATH=fgnn-artifacts_b200 python examples/train_graphsage_multi_gpu.py --dataset-path /data/papers100M \\
      --num-sample-worker 2 --num-train-worker 6 --cache-percentage 0.25 --num-epoch 4 --pipeline

--model gcn | pinsage selects the reference's other two model families (examples/gnn_models_csc.py).
--no-train skips the model (the trainers only take the batches and read the labels back): the path the bench's
`e2e_factored` leg times.  --json prints one machine-readable line with per-epoch times (bench.py extra.epoch).
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "fgnn-artifacts_b200"))
sys.path.insert(0, HERE)


def parse(argv=None):
    ap = argparse.ArgumentParser("factored GraphSAGE on samgraph-b200 (DGL-free)")
    ap.add_argument("--dataset-path", required=True)
    ap.add_argument("--num-sample-worker", type=int, default=1)
    ap.add_argument("--num-train-worker", type=int, default=1)
    ap.add_argument("--single-gpu", action="store_true", help="samplers and trainers all on cuda:0")
    ap.add_argument("--sampler-device", default=None,
                    help="put every sampler on this device (e.g. cuda:0) instead of cuda:T.. : lets a box with T GPUs "
                         "run T trainers with DDP plus a sampler that shares a trainer's GPU")
    ap.add_argument("--sample-type", default="khop2")
    ap.add_argument("--fanout", nargs="+", type=int, default=[25, 10])
    ap.add_argument("--batch-size", type=int, default=8000)
    ap.add_argument("--num-epoch", type=int, default=4)
    ap.add_argument("--num-hidden", type=int, default=256)
    ap.add_argument("--model", default="graphsage", choices=["graphsage", "gcn", "pinsage"],
                    help="trainer-side consumer: the reference's multi_gpu/train_graphsage.py, train_gcn.py or "
                         "train_pinsage.py model (examples/gnn_models_csc.py); pinsage needs --sample-type random_walk")
    ap.add_argument("--lr", type=float, default=0.003)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--cache-policy", default="pre_sample")
    ap.add_argument("--cache-percentage", type=float, default=0.25)
    ap.add_argument("--replicate-percentage", type=float, default=None)
    ap.add_argument("--no-partition-cache", action="store_true")
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-ddp", action="store_true")
    ap.add_argument("--max-copying-jobs", type=int, default=4)
    ap.add_argument("--seed", type=int, default=0x5EED)
    ap.add_argument("--master-port", type=int, default=12377)
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--timeout", type=int, default=300)
    return ap.parse_args(argv)


def run_config(a, sam):
    cfg = {"dataset_path": a.dataset_path, "arch": "arch5", "_arch": sam.builtin_archs["arch5"]["arch"],
           "sample_type": a.sample_type, "_sample_type": sam.sample_types[a.sample_type],
           "batch_size": a.batch_size, "num_epoch": a.num_epoch, "cache_policy": a.cache_policy,
           "_cache_policy": sam.cache_policies[a.cache_policy], "cache_percentage": a.cache_percentage,
           "max_sampling_jobs": 10, "max_copying_jobs": a.max_copying_jobs, "omp_thread_num": os.cpu_count() or 1,
           "num_sample_worker": a.num_sample_worker, "num_train_worker": a.num_train_worker, "presample_epoch": 1,
           "seed": a.seed, "partition_cache": 0 if a.no_partition_cache else 1}
    if a.replicate_percentage is not None:
        cfg["replicate_percentage"] = a.replicate_percentage
    if a.sample_type == "random_walk":
        cfg.update(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5, num_layer=3)
    else:
        cfg.update(fanout=a.fanout, num_fanout=len(a.fanout), num_layer=len(a.fanout))
    return cfg


def _dump_on_sigusr1():
    """a hung worker prints its Python stacks when the parent gives up (SIGUSR1), before it is killed"""
    import faulthandler
    import signal
    faulthandler.register(signal.SIGUSR1, all_threads=True)


def run_sample(worker_id, a, ctx, barrier, outdir):
    """train_graphsage.py:95-215"""
    _dump_on_sigusr1()
    import samgraph.torch as sam
    sam.sample_init(worker_id, ctx)
    sam.notify_sampler_ready(barrier)
    num_epoch, num_step = sam.num_epoch(), sam.num_local_step()
    barrier.wait()                                         # run start
    times = []
    for epoch in range(num_epoch):
        barrier.wait()                                     # epoch start
        tic = time.time()
        for _ in range(num_step):
            sam.sample_once()
        times.append(time.time() - tic)
        barrier.wait()                                     # epoch end
    barrier.wait()                                         # run end
    out = {"role": "sampler", "worker": worker_id, "ctx": ctx, "local_steps": num_step, "epoch_wall_s": times,
           "epoch_sample_s": [sam.get_log_epoch_value(e, sam.kLogEpochSampleTime) for e in range(num_epoch)],
           "epoch_send_s": [sam.get_log_epoch_value(e, sam.kLogEpochSampleSendTime) for e in range(num_epoch)],
           "init_s": sam.get_log_init_value(sam.kLogInitL1Sampler)}
    json.dump(out, open(os.path.join(outdir, "s%d.json" % worker_id), "w"))
    sam.shutdown()


def run_train(worker_id, a, ctx, barrier, outdir):
    """train_graphsage.py:217-437"""
    _dump_on_sigusr1()
    import torch
    import samgraph.torch as sam
    T = a.num_train_worker
    sam.wait_for_sampler_ready(barrier)
    sam.train_init(worker_id, ctx)
    dev = torch.device(ctx)
    torch.cuda.set_device(dev)
    L = 3 if a.sample_type == "random_walk" else len(a.fanout)
    model = None
    if not a.no_train:
        import torch.nn as nn
        from train_graphsage_csc import SAGE, csc_blocks
        if T > 1 and not a.no_ddp:
            for k in [k for k in os.environ if k.startswith("TORCHELASTIC_")]:
                os.environ.pop(k)          # not a torchrun worker even when launched from one (agent store!)
            torch.distributed.init_process_group(backend="nccl", init_method="tcp://127.0.0.1:%d" % a.master_port,
                                                 world_size=T, rank=worker_id, device_id=dev)
        if a.model == "graphsage":
            model = SAGE(sam.feat_dim(), a.num_hidden, sam.num_class(), L, a.dropout).to(dev)
        else:
            from gnn_models_csc import build_model, csc_blocks_weighted
            assert a.model != "pinsage" or a.sample_type == "random_walk", "pinsage consumes the random-walk edge weights"
            model = build_model(a.model, sam.feat_dim(), a.num_hidden, sam.num_class(), L, a.dropout).to(dev)
            if a.model == "pinsage":
                csc_blocks = csc_blocks_weighted
        if T > 1 and not a.no_ddp:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev], output_device=dev)
        loss_fn = nn.CrossEntropyLoss().to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=a.lr)
        model.train()
    num_epoch, num_step = sam.num_epoch(), sam.steps_per_epoch()
    align_up = (num_step + T - 1) // T * T
    barrier.wait()                                         # run start
    ep_wall, ep_edges, ep_rows, ep_steps, losses = [], [], [], [], []
    for epoch in range(num_epoch):
        barrier.wait()                                     # epoch start
        tic = time.time()
        need = num_step // T + (1 if worker_id < num_step % T else 0)
        if a.pipeline:
            sam.extract_start(need)
        edges = rows = steps = 0
        loss = None
        blocks = feat = label = None
        for step in range(worker_id, align_up, T):
            if step < num_step:
                t0 = time.time()
                if not a.pipeline:
                    sam.sample_once()
                key = sam.get_next_batch()
                t1 = time.time()
                if model is not None:
                    blocks, feat, label = csc_blocks(sam, key, L)
                else:
                    label = sam.get_graph_label(key)
                    feat = sam.get_graph_feat(key)
                t2 = time.time()
                for i in range(L):
                    edges += sam.get_graph_num_edge(key, i)
                rows += feat.shape[0]
                steps += 1
            if model is not None and blocks is not None:
                loss = loss_fn(model(blocks, feat), label)
                opt.zero_grad()
                loss.backward()
                opt.step()
                torch.cuda.synchronize()                   # event_sync(): the batch may be freed after this
            elif label is not None and step < num_step:
                label.cpu()                                # device -> host read of the step's result
            if step + T < num_step:                        # the last batch stays: trainers with one step less
                blocks = feat = label = None               # repeat it so that DDP's collectives stay aligned (:298-323)
            if step < num_step:
                t3 = time.time()
                sam.log_step(epoch, step, sam.kLogL1ConvertTime, t2 - t1)
                sam.log_step(epoch, step, sam.kLogL1TrainTime, t3 - t2)
                sam.log_epoch_add(epoch, sam.kLogEpochConvertTime, t2 - t1)
                sam.log_epoch_add(epoch, sam.kLogEpochTrainTime, t3 - t2)
                sam.log_epoch_add(epoch, sam.kLogEpochTotalTime, t3 - t0)
        torch.cuda.synchronize()
        ep_wall.append(time.time() - tic)
        ep_edges.append(edges)
        ep_rows.append(rows)
        ep_steps.append(steps)
        losses.append(float(loss.detach()) if loss is not None else None)
        barrier.wait()                                     # epoch end
    barrier.wait()                                         # run end
    ge = sam.get_log_epoch_value
    out = {"role": "trainer", "worker": worker_id, "ctx": ctx, "epoch_wall_s": ep_wall, "epoch_edges": ep_edges,
           "epoch_rows": ep_rows, "epoch_steps": ep_steps, "loss": losses,
           "epoch_copy_s": [ge(e, sam.kLogEpochCopyTime) for e in range(num_epoch)],
           "epoch_convert_s": [ge(e, sam.kLogEpochConvertTime) for e in range(num_epoch)],
           "epoch_train_s": [ge(e, sam.kLogEpochTrainTime) for e in range(num_epoch)],
           "epoch_feature_bytes": [ge(e, sam.kLogEpochFeatureBytes) for e in range(num_epoch)],
           "epoch_miss_bytes": [ge(e, sam.kLogEpochMissBytes) for e in range(num_epoch)],
           "init_s": sam.get_log_init_value(sam.kLogInitL1Trainer)}
    json.dump(out, open(os.path.join(outdir, "t%d.json" % worker_id), "w"))
    sam.shutdown()


def main(argv=None):
    a = parse(argv)
    import samgraph.torch as sam
    S, T = a.num_sample_worker, a.num_train_worker
    if a.single_gpu and T > 1:
        a.no_ddp = True      # NCCL refuses two ranks on one device; the reference forces T = 1 there (common_config.py:186-191)
    sam.config(run_config(a, sam))
    sam.data_init()                                        # no CUDA before fork (dist_engine.cc:611-632)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(S + T, timeout=a.timeout)
    t_ctx = ["cuda:0"] * T if a.single_gpu else ["cuda:%d" % i for i in range(T)]
    s_ctx = ["cuda:0"] * S if a.single_gpu else ["cuda:%d" % (T + i) for i in range(S)]
    if a.sampler_device:
        s_ctx = [a.sampler_device] * S
    outdir = tempfile.mkdtemp(prefix="fgnn_factored_")
    t_start = time.time()
    procs = [ctx.Process(target=run_sample, args=(i, a, s_ctx[i], barrier, outdir)) for i in range(S)] + \
            [ctx.Process(target=run_train, args=(i, a, t_ctx[i], barrier, outdir)) for i in range(T)]
    for p in procs:
        p.start()
    bad = 0
    deadline = time.time() + a.timeout
    # like sam.wait_one_child() in the reference's scripts: the first worker that dies takes the job down, the others
    # must not sit in a barrier until it times out
    while any(p.is_alive() for p in procs):
        timed_out = time.time() > deadline
        if any(p.exitcode not in (None, 0) for p in procs) or timed_out:
            if timed_out:
                import signal
                print("train_graphsage_multi_gpu: timeout after %d s, worker stacks follow" % a.timeout, file=sys.stderr)
                for p in procs:
                    if p.is_alive():
                        os.kill(p.pid, signal.SIGUSR1)
                time.sleep(1.0)
            for p in procs:
                if p.is_alive():
                    p.kill()
            break
        time.sleep(0.05)
    for p in procs:
        p.join(10)
        bad |= (p.exitcode != 0)
    if bad:
        print("train_graphsage_multi_gpu: a worker failed", file=sys.stderr)
        sys.exit(1)
    res = [json.load(open(os.path.join(outdir, f))) for f in sorted(os.listdir(outdir))]
    trainers = [r for r in res if r["role"] == "trainer"]
    samplers = [r for r in res if r["role"] == "sampler"]
    E = a.num_epoch
    epochs = []
    for e in range(E):
        epochs.append({"wall_s": max(t["epoch_wall_s"][e] for t in trainers),
                       "edges": sum(t["epoch_edges"][e] for t in trainers),
                       "steps": sum(t["epoch_steps"][e] for t in trainers),
                       "rows": sum(t["epoch_rows"][e] for t in trainers),
                       "sample_s": max(s["epoch_sample_s"][e] for s in samplers),
                       "sampler_wall_s": max(s["epoch_wall_s"][e] for s in samplers),
                       "copy_s": max(t["epoch_copy_s"][e] for t in trainers),
                       "convert_s": max(t["epoch_convert_s"][e] for t in trainers),
                       "train_s": max(t["epoch_train_s"][e] for t in trainers),
                       "miss_bytes": sum(t["epoch_miss_bytes"][e] for t in trainers),
                       "feature_bytes": sum(t["epoch_feature_bytes"][e] for t in trainers),
                       "per_trainer_wall_s": [t["epoch_wall_s"][e] for t in trainers]})
    for e, ep in enumerate(epochs):
        print("Epoch {:03d} | time {:.4f} s | sample {:.4f} | copy {:.4f} | convert {:.4f} | train {:.4f} | edges {:d}".format(
            e, ep["wall_s"], ep["sample_s"], ep["copy_s"], ep["convert_s"], ep["train_s"], ep["edges"]))
    timed = epochs[1:] if E > 1 else epochs                # the scripts drop the warm-up epoch (common_config.py:163)
    avg = sum(ep["wall_s"] for ep in timed) / len(timed)
    print("test_result:epoch_time:total={:.4f}".format(avg))
    if a.json:
        print("FACTORED_JSON " + json.dumps({
            "samplers": S, "trainers": T, "sampler_ctx": s_ctx, "trainer_ctx": t_ctx, "pipeline": a.pipeline,
            "train": not a.no_train, "cache_percentage": a.cache_percentage, "epochs": epochs,
            "avg_epoch_s": avg, "edges_per_s": sum(ep["edges"] for ep in timed) / sum(ep["wall_s"] for ep in timed),
            "steps_timed": sum(ep["steps"] for ep in timed), "init_s": time.time() - t_start - sum(ep["wall_s"] for ep in epochs),
            "loss": [t["loss"] for t in trainers][0]}))


if __name__ == "__main__":
    main()
