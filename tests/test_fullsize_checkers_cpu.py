"""The property checkers of tests/test_zz_fullsize_configs_gpu.py, run on CPU tensors holding ORACLE batches of a
small graph: whatever the full-size GPU tests assert must hold for the reference semantics (and a corrupted batch
must be caught), otherwise a GPU failure there would say nothing."""
import importlib.util
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def chk():
    spec = importlib.util.spec_from_file_location("fullsize_cfg", os.path.join(ROOT, "tests", "test_zz_fullsize_configs_gpu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def t64(a):
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int64))


def as_big(indptr, indices):
    return dict(indptr=torch.from_numpy(indptr.view(np.int32)), indices=torch.from_numpy(indices.view(np.int32)))


def test_checkers_accept_oracle_batches_and_reject_corruption(chk, oracle, graph_medium):
    from oracle.oracle import sample_batch_oracle
    indptr, indices = graph_medium
    V = len(indptr) - 1
    big = as_big(indptr, indices)
    rng = np.random.default_rng(3)
    seeds = rng.permutation(V)[:600].astype(np.uint32)
    g = dict(indptr=indptr, indices=indices)

    # --- uniform (GCN-like, three layers) ---
    fan = [3, 4, 5]
    exp = sample_batch_oracle(oracle, g, seeds, fan, "khop2", 9, 2)
    n2o = t64(exp["input_nodes"])
    for i, e in enumerate(exp["layers"]):
        chk.check_uniform_layer(big, n2o, t64(e["row"]), t64(e["col"]), e["num_dst"], e["num_src"], fan[i])
    e = exp["layers"][0]
    bad = t64(e["row"]).clone()
    bad[5] = (bad[5] + 1) % e["num_src"]                   # an edge that is (almost surely) not in the graph
    with pytest.raises(AssertionError):
        chk.check_uniform_layer(big, n2o, bad, t64(e["col"]), e["num_dst"], e["num_src"], fan[0])

    # --- PinSAGE random walk ---
    rw = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=5)
    exp = sample_batch_oracle(oracle, g, seeds, [5, 5, 5], "random_walk", 9, 2, rw)
    n2o = t64(exp["input_nodes"])
    for e in exp["layers"]:
        chk.check_random_walk_layer(big, n2o, t64(e["row"]), t64(e["col"]), t64(e["data"]), e["num_dst"], e["num_src"],
                                    5, 12)
    e = exp["layers"][2]
    bad = t64(e["data"]).clone()
    bad[0] = 13                                            # more visits than 4 walks x 3 steps can record
    with pytest.raises(AssertionError):
        chk.check_random_walk_layer(big, n2o, t64(e["row"]), t64(e["col"]), bad, e["num_dst"], e["num_src"], 5, 12)

    # --- weighted k-hop (alias) ---
    w = rng.integers(1, 11, size=len(indices)).astype(np.float32)
    prob, alias = oracle.build_alias_table(indptr, indices, w)
    gw = dict(g, prob_table=prob, alias_table=alias)
    fan = [6, 4]
    exp = sample_batch_oracle(oracle, gw, seeds, fan, "weighted_khop", 9, 2)
    n2o = t64(exp["input_nodes"])
    for i, e in enumerate(exp["layers"]):
        chk.check_weighted_layer(big, n2o, t64(e["row"]), t64(e["col"]), e["num_dst"], e["num_src"], fan[i])
    e = exp["layers"][1]
    col = t64(e["col"])
    with pytest.raises(AssertionError):                    # seed order destroyed
        chk.check_weighted_layer(big, n2o, t64(e["row"]), col.flip(0), e["num_dst"], e["num_src"], fan[1])
