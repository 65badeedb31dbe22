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


def load_models():
    spec = importlib.util.spec_from_file_location("gnn_models_csc", os.path.join(ROOT, "examples", "gnn_models_csc.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def oracle_blocks(oracle, graph_small, sample_type, fanouts, rw=None, seed=7):
    from oracle.oracle import sample_batch_oracle
    indptr, indices = graph_small
    V = len(indptr) - 1
    rng = np.random.default_rng(seed)
    seeds = rng.permutation(V)[:150].astype(np.uint32)
    exp = sample_batch_oracle(oracle, dict(indptr=indptr, indices=indices), seeds, fanouts, sample_type, 5, 1, rw=rw)
    csc, coo = [], []
    for e in exp["layers"]:
        ip, idx, eids = oracle.coo_to_csc(e["row"], e["col"], e["num_dst"])
        w = None
        if e["data"] is not None:
            w = torch.from_numpy(e["data"][eids].astype(np.int32))          # CSC edge order
        csc.append((torch.from_numpy(ip.astype(np.int32)), torch.from_numpy(idx.astype(np.int32)), e["num_src"],
                    e["num_dst"]) + ((w,) if w is not None else ()))
        coo.append((torch.from_numpy(e["row"].astype(np.int64)), torch.from_numpy(e["col"].astype(np.int64)),
                    e["num_src"], e["num_dst"],
                    None if e["data"] is None else torch.from_numpy(e["data"].astype(np.float32))))
    return exp, seeds, csc, coo, rng


def test_gcn_csc_matches_graphconv_definition(oracle, graph_small):
    """dgl GraphConv(norm='both', allow_zero_in_degree=True) (train_gcn.py:18-47) written with index_add on the COO
    blocks vs the CSC SpMM layer, forward and feature gradient; GCN [5,10,15] has three layers."""
    M = load_models()
    fanouts = [3, 4, 5]
    exp, seeds, csc, coo, rng = oracle_blocks(oracle, graph_small, "khop2", fanouts)
    D, H, C = 10, 14, 6            # in > out on the last layer: the multiply-first branch is exercised too
    feat = torch.from_numpy(rng.standard_normal((len(exp["input_nodes"]), D)).astype(np.float32)).requires_grad_(True)
    torch.manual_seed(1)
    model = M.GCN(D, H, C, len(fanouts), dropout=0.0)
    out = model(csc, feat)
    assert out.shape == (len(seeds), C)

    def reference(h):
        for i, (row, col, num_src, num_dst, _) in enumerate(coo):
            layer = model.layers[i]
            out_deg = torch.zeros(num_src).index_add_(0, row, torch.ones(len(row))).clamp(min=1)
            in_deg = torch.zeros(num_dst).index_add_(0, col, torch.ones(len(col))).clamp(min=1)
            hs = h * out_deg.pow(-0.5)[:, None]
            agg = torch.zeros((num_dst, hs.shape[1])).index_add_(0, col, hs[row])
            h = (agg @ layer.weight) * in_deg.pow(-0.5)[:, None] + layer.bias
            if i != len(coo) - 1:
                h = torch.relu(h)
        return h

    ref = reference(feat)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)
    g1, = torch.autograd.grad(out.square().sum(), feat, retain_graph=True)
    g2, = torch.autograd.grad(ref.square().sum(), feat)
    assert torch.allclose(g1, g2, rtol=1e-3, atol=1e-5)


def test_pinsage_csc_matches_weighted_sage_conv_definition(oracle, graph_small):
    """WeightedSAGEConv (train_pinsage.py:30-66) with the random-walk sampler's visit counts as edge weights:
    index_add formulation on the COO blocks vs the CSC SpMM layer; rows come out with unit L2 norm."""
    M = load_models()
    rw = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5)
    exp, seeds, csc, coo, rng = oracle_blocks(oracle, graph_small, "random_walk", [5, 5, 5], rw=rw)
    assert all(len(b) == 5 for b in csc)
    D, H, C = 9, 12, 7
    feat = torch.from_numpy(rng.standard_normal((len(exp["input_nodes"]), D)).astype(np.float32)).requires_grad_(True)
    torch.manual_seed(2)
    model = M.PinSAGE(D, H, C, 3, dropout=0.0)
    out = model(csc, feat)
    assert out.shape == (len(seeds), C)
    norms = out.norm(2, 1)
    assert torch.all((norms - 1).abs() < 1e-5) or torch.all((norms < 1e-5) | ((norms - 1).abs() < 1e-5))

    def reference(h):
        for i, (row, col, num_src, num_dst, w) in enumerate(coo):
            layer = model.layers[i]
            n = torch.relu(layer.Q(h))
            agg = torch.zeros((num_dst, n.shape[1])).index_add_(0, col, n[row] * w[:, None])
            ws = torch.zeros(num_dst).index_add_(0, col, w).clamp(min=1)
            z = torch.relu(layer.W(torch.cat([agg / ws[:, None], h[:num_dst]], 1)))
            zn = z.norm(2, 1, keepdim=True)
            h = z / torch.where(zn == 0, torch.ones_like(zn), zn)
        return h

    ref = reference(feat)
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5)
    g1, = torch.autograd.grad(out[:, 0].sum(), feat, retain_graph=True)
    g2, = torch.autograd.grad(ref[:, 0].sum(), feat)
    assert torch.allclose(g1, g2, rtol=1e-3, atol=1e-5)
