#!/usr/bin/env python
"""GraphSAGE training on the samgraph B200 runtime without DGL / PyG: the loop of the reference's
example/samgraph/train_graphsage.py (same sam.* calls, same per-epoch report), with the blocks taken in CSC form
(sam.get_csc_blocks — the trainer pays no COO->CSC conversion) and the mean aggregation written as one sparse-CSR
SpMM per layer.  The reference scripts themselves run unchanged on this runtime once DGL is installed; this file is
for boxes where it is not.

  PYTHONPATH=fgnn-artifacts_b200 python examples/train_graphsage_csc.py --dataset-path /data/papers100M \\
      --cache-percentage 0.25 --num-epoch 3 [--pipeline]
  PYTHONPATH=fgnn-artifacts_b200 python examples/train_graphsage_csc.py --synthetic ci-1m      # generated dataset

STATUS: the model below is unit-tested on CPU (tests/test_example_model_cpu.py) and trains on B200s as the consumer of
examples/train_graphsage_multi_gpu.py (bench.py's measured-epoch leg, tests/test_runtime_gpu.py); the single-process
main() of THIS file has not been run on a GPU.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn
import torch.nn.functional as F


class SAGEConvCSC(nn.Module):
    """h_dst' = W_self h_dst + W_neigh mean_{src in N(dst)} h_src + b, with the block given as CSC of the
    (src -> dst) bipartite graph: indptr over dst nodes, indices = src local ids (dst nodes are a prefix of src)."""

    def __init__(self, in_feats, out_feats):
        super().__init__()
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=True)

    def forward(self, block, h):
        indptr, indices, num_src, num_dst = block
        assert h.shape[0] == num_src
        deg = (indptr[1:] - indptr[:-1]).to(h.dtype).clamp(min=1)
        # row d of A holds 1/deg(d) at the columns of d's sampled neighbours: mean aggregation = A @ h
        vals = torch.repeat_interleave(1.0 / deg, (indptr[1:] - indptr[:-1]).long())
        adj = torch.sparse_csr_tensor(indptr, indices, vals, size=(num_dst, num_src))
        mean = torch.sparse.mm(adj, h)
        return self.fc_self(h[:num_dst]) + self.fc_neigh(mean)


class SAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, dropout):
        super().__init__()
        dims = [in_feats] + [n_hidden] * (n_layers - 1) + [n_classes]
        self.layers = nn.ModuleList(SAGEConvCSC(dims[i], dims[i + 1]) for i in range(n_layers))
        self.dropout = nn.Dropout(dropout)

    def forward(self, blocks, x):
        h = x
        for i, (layer, block) in enumerate(zip(self.layers, blocks)):
            h = layer(block, h)
            if i != len(self.layers) - 1:
                h = self.dropout(F.relu(h))
        return h


def csc_blocks(sam, batch_key, num_layers):
    """[(indptr, indices, num_src, num_dst)] per layer, input side first, as int64-free device tensors."""
    blocks, feat, label = sam.get_csc_blocks(batch_key, num_layers)
    return [(indptr, indices, num_src, num_dst) for indptr, indices, _eids, num_src, num_dst in blocks], feat, label


def parse():
    ap = argparse.ArgumentParser("GraphSAGE on samgraph-b200 (DGL-free)")
    ap.add_argument("--dataset-path", default=None)
    ap.add_argument("--synthetic", default=None, help="generate a synthetic dataset of this shape (fgnn_b200.synth.SHAPES)")
    ap.add_argument("--arch", default="arch3", choices=["arch1", "arch2", "arch3"])
    ap.add_argument("--sample-type", default="khop2")
    ap.add_argument("--fanout", nargs="+", type=int, default=[25, 10])
    ap.add_argument("--batch-size", type=int, default=8000)
    ap.add_argument("--num-epoch", type=int, default=3)
    ap.add_argument("--num-hidden", type=int, default=256)
    ap.add_argument("--model", default="graphsage", choices=["graphsage", "gcn", "pinsage"],
                    help="gcn / pinsage: the reference's train_gcn.py / train_pinsage.py models (examples/gnn_models_csc.py)")
    ap.add_argument("--lr", type=float, default=0.003)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--cache-policy", default="pre_sample")
    ap.add_argument("--cache-percentage", type=float, default=0.25)
    ap.add_argument("--pipeline", action="store_true")
    ap.add_argument("--device", default="cuda:0")
    return ap.parse_args()


def main():
    a = parse()
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "fgnn-artifacts_b200"))
    import samgraph.torch as sam
    path = a.dataset_path
    if a.synthetic:
        from fgnn_b200.synth import make_dataset_numpy, write_dataset
        path = "/dev/shm/fgnn_example_%s" % a.synthetic
        if not os.path.exists(os.path.join(path, "meta.txt")):
            write_dataset(path, make_dataset_numpy(a.synthetic))
    assert path, "--dataset-path or --synthetic is required"
    cfg = {"dataset_path": path, "arch": a.arch, "_arch": sam.builtin_archs[a.arch]["arch"],
           "sample_type": a.sample_type, "_sample_type": sam.sample_types[a.sample_type],
           "batch_size": a.batch_size, "num_epoch": a.num_epoch, "cache_policy": a.cache_policy,
           "_cache_policy": sam.cache_policies[a.cache_policy], "cache_percentage": a.cache_percentage,
           "max_sampling_jobs": 10, "max_copying_jobs": 2, "omp_thread_num": os.cpu_count() or 1,
           "sampler_ctx": a.device, "trainer_ctx": a.device, "fanout": a.fanout, "num_fanout": len(a.fanout),
           "num_layer": len(a.fanout), "presample_epoch": 1}
    if a.sample_type == "random_walk":         # train_pinsage.py:122-126
        for k in ("fanout", "num_fanout"):
            cfg.pop(k)
        cfg.update(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5, num_layer=3)
    sam.config(cfg)
    sam.init()
    dev = torch.device(a.device)
    torch.cuda.set_device(dev)
    L = 3 if a.sample_type == "random_walk" else len(a.fanout)
    get_blocks = csc_blocks
    if a.model == "graphsage":
        model = SAGE(sam.feat_dim(), a.num_hidden, sam.num_class(), L, a.dropout).to(dev)
    else:
        sys.path.insert(0, here)
        from gnn_models_csc import build_model, csc_blocks_weighted
        assert a.model != "pinsage" or a.sample_type == "random_walk", "pinsage consumes the random-walk edge weights"
        model = build_model(a.model, sam.feat_dim(), a.num_hidden, sam.num_class(), L, a.dropout).to(dev)
        if a.model == "pinsage":
            get_blocks = csc_blocks_weighted
    loss_fn = nn.CrossEntropyLoss()
    opt = torch.optim.Adam(model.parameters(), lr=a.lr)
    num_epoch, num_step = sam.num_epoch(), sam.steps_per_epoch()
    model.train()
    if a.pipeline:
        sam.start()
    totals = []
    for epoch in range(num_epoch):
        t_epoch = time.time()
        for step in range(num_step):
            t0 = time.time()
            if not a.pipeline:
                sam.sample_once()
            batch_key = sam.get_next_batch()
            t1 = time.time()
            blocks, feat, label = get_blocks(sam, batch_key, L)
            t2 = time.time()
            loss = loss_fn(model(blocks, feat), label)
            opt.zero_grad()
            loss.backward()
            opt.step()
            torch.cuda.synchronize()
            t3 = time.time()
            sam.log_step(epoch, step, sam.kLogL1ConvertTime, t2 - t1)
            sam.log_step(epoch, step, sam.kLogL1TrainTime, t3 - t2)
            sam.log_epoch_add(epoch, sam.kLogEpochConvertTime, t2 - t1)
            sam.log_epoch_add(epoch, sam.kLogEpochTrainTime, t3 - t2)
            sam.log_epoch_add(epoch, sam.kLogEpochTotalTime, t3 - t0)
        totals.append(time.time() - t_epoch)
        print("Epoch {:03d} | time {:.4f} s | sample {:.4f} | copy {:.4f} | convert {:.4f} | train {:.4f} | loss {:.4f}".format(
            epoch, totals[-1], sam.get_log_epoch_value(epoch, sam.kLogEpochSampleTime),
            sam.get_log_epoch_value(epoch, sam.kLogEpochCopyTime), sam.get_log_epoch_value(epoch, sam.kLogEpochConvertTime),
            sam.get_log_epoch_value(epoch, sam.kLogEpochTrainTime), float(loss)))
        sam.forward_barrier()
    print("test_result:epoch_time:total={:.4f}".format(sum(totals[1:]) / max(1, len(totals) - 1)))
    sam.report_step_average(num_epoch - 1, num_step - 1)
    sam.shutdown()


if __name__ == "__main__":
    main()
