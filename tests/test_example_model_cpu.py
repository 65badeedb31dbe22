"""examples/train_graphsage_csc.py: the DGL-free GraphSAGE layer (mean aggregation as one sparse-CSR SpMM over the
CSC hand-off) against a plain index_add formulation on oracle-produced blocks, forward and backward, on CPU."""
import importlib.util
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_example():
    spec = importlib.util.spec_from_file_location("train_graphsage_csc", os.path.join(ROOT, "examples", "train_graphsage_csc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_sage_csc_layers_match_index_add_reference(oracle, graph_small):
    from oracle.oracle import sample_batch_oracle
    ex = load_example()
    indptr, indices = graph_small
    V = len(indptr) - 1
    rng = np.random.default_rng(0)
    seeds = rng.permutation(V)[:200].astype(np.uint32)
    fanouts = [6, 4]
    exp = sample_batch_oracle(oracle, dict(indptr=indptr, indices=indices), seeds, fanouts, "khop2", 5, 1)
    D, H, C = 12, 16, 5
    feat = torch.from_numpy(rng.standard_normal((len(exp["input_nodes"]), D)).astype(np.float32)).requires_grad_(True)
    blocks_csc, blocks_coo = [], []
    for e in exp["layers"]:
        ip, idx, eids = oracle.coo_to_csc(e["row"], e["col"], e["num_dst"])
        blocks_csc.append((torch.from_numpy(ip.astype(np.int32)), torch.from_numpy(idx.astype(np.int32)), e["num_src"], e["num_dst"]))
        blocks_coo.append((torch.from_numpy(e["row"].astype(np.int64)), torch.from_numpy(e["col"].astype(np.int64)),
                           e["num_src"], e["num_dst"]))
    torch.manual_seed(0)
    model = ex.SAGE(D, H, C, len(fanouts), dropout=0.0)
    out = model(blocks_csc, feat)
    assert out.shape == (len(seeds), C)

    def reference(h):
        for i, (row, col, num_src, num_dst) in enumerate(blocks_coo):
            layer = model.layers[i]
            agg = torch.zeros((num_dst, h.shape[1])).index_add_(0, col, h[row])
            deg = torch.zeros(num_dst).index_add_(0, col, torch.ones(len(col))).clamp(min=1)
            h = layer.fc_self(h[:num_dst]) + layer.fc_neigh(agg / deg[:, None])
            if i != len(blocks_coo) - 1:
                h = torch.relu(h)
        return h

    ref = reference(feat)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    # backward through the SpMM: same feature gradient as the dense formulation
    g1, = torch.autograd.grad(out.square().sum(), feat, retain_graph=True)
    g2, = torch.autograd.grad(ref.square().sum(), feat)
    assert torch.allclose(g1, g2, rtol=1e-4, atol=1e-5)
