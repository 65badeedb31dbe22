"""GPU parity tests of the dataset-preparation kernels: alias / prefix weight tables, out-degree
and the cache_by_degree / cache_by_random rankings, against the CPU oracle (bit-exact, fp32 words compared as u32)."""
import numpy as np
import pytest

from test_kernels_gpu import K, G, dev, gm, gs, host  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def f32_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def weights_for(indices, kind, seed=2):
    rng = np.random.default_rng(seed)
    if kind == "int1_10":                       # the reference's kDefault policy (create_alias_table.cc:113)
        return rng.integers(1, 11, size=len(indices)).astype(np.float32)
    if kind == "uniform":
        return (rng.random(len(indices), dtype=np.float32) + np.float32(1e-3)).astype(np.float32)
    if kind == "skewed":                        # a few dominant edges per row: long large->small chains
        w = rng.random(len(indices), dtype=np.float32) ** 8 * np.float32(1000.0) + np.float32(0.01)
        return w.astype(np.float32)
    return np.ones(len(indices), np.float32)    # all weights equal: every scaled weight is exactly 1 -> all "large"


@pytest.mark.parametrize("graph", ["small", "medium", "hub"])
@pytest.mark.parametrize("kind", ["int1_10", "uniform", "skewed", "ones"])
def test_alias_and_prefix_tables_match_oracle(K, oracle, gs, gm, graph, kind):
    if graph == "hub":
        # one row with 200k neighbours, one with a single neighbour, isolated vertices in between
        deg = np.zeros(64, np.int64)
        deg[3], deg[10], deg[40:50] = 200000, 1, 17
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
        indices = np.random.default_rng(5).integers(0, 64, size=int(deg.sum())).astype(np.uint32)
        g = G(indptr, indices)
    else:
        g = gs if graph == "small" else gm
    V, E = len(g.indptr_np) - 1, len(g.indices_np)
    w = weights_for(g.indices_np, kind)
    exp_prob, exp_alias = oracle.build_alias_table(g.indptr_np, g.indices_np, w)
    exp_prefix = oracle.build_prefix_table(g.indptr_np, w)
    d_w = f32_dev(w)
    prob = torch.full((E,), -7.0, dtype=torch.float32, device="cuda")
    alias = torch.full((E,), -1, dtype=torch.int32, device="cuda")
    prefix = torch.full((E,), -7.0, dtype=torch.float32, device="cuda")
    for _ in range(2):                          # twice: the row ticket must be re-armed by the first launch
        K.build_alias_table(g.indptr, g.indices, V, E, d_w, prob, alias)
        K.build_prefix_table(g.indptr, V, d_w, prefix)
    torch.cuda.synchronize()
    assert np.array_equal(prob.cpu().numpy().view(np.uint32), exp_prob.view(np.uint32))
    assert np.array_equal(host(alias), exp_alias)
    assert np.array_equal(prefix.cpu().numpy().view(np.uint32), exp_prefix.view(np.uint32))
    # the tables feed the alias sampler unchanged: every prob in (0, 1], every alias a neighbour of its row
    p = prob.cpu().numpy()
    assert np.all((p > 0) & (p <= 1))


def test_alias_table_empty_graph(K):
    indptr = dev(np.zeros(5, np.uint32))
    e = torch.empty(0, dtype=torch.int32, device="cuda")
    K.build_alias_table(indptr, e, 4, 0, e.float(), e.float(), e, workspace=torch.empty(1, dtype=torch.uint8, device="cuda"))
    K.build_prefix_table(indptr, 4, torch.empty(1, device="cuda"), torch.empty(1, device="cuda"))
    torch.cuda.synchronize()


@pytest.mark.parametrize("graph", ["small", "medium"])
def test_out_degree_and_degree_ranking(K, oracle, gs, gm, graph):
    g = gs if graph == "small" else gm
    V, E = len(g.indptr_np) - 1, len(g.indices_np)
    exp_deg = np.bincount(g.indices_np, minlength=V).astype(np.uint32)      # graph_loader.cc:126-137
    deg = torch.empty(V, dtype=torch.int32, device="cuda")
    rank = torch.empty(V, dtype=torch.int32, device="cuda")
    ws = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device="cuda")
    # unaligned view of indices exercises the scalar head/tail of the vectorised histogram
    K.out_degree(g.indices[1:], E - 1, deg, V)
    torch.cuda.synchronize()
    assert np.array_equal(host(deg), np.bincount(g.indices_np[1:], minlength=V).astype(np.uint32))
    K.rank_by_degree(g.indices, E, V, deg, rank, ws)
    torch.cuda.synchronize()
    assert np.array_equal(host(deg), exp_deg)
    # cache_by_degree.cc:36-47: std::greater on pair{out_degree, id} == the PreSC order with freq := degree
    assert np.array_equal(host(rank), oracle.presc_rank(exp_deg))
    order = np.lexsort((-np.arange(V, dtype=np.int64), -exp_deg.astype(np.int64)))
    assert np.array_equal(host(rank), order.astype(np.uint32))


@pytest.mark.parametrize("V", [1, 1000, 300001])
def test_random_ranking_is_a_seeded_permutation(K, V):
    a = torch.empty(V, dtype=torch.int32, device="cuda")
    b = torch.empty(V, dtype=torch.int32, device="cuda")
    c = torch.empty(V, dtype=torch.int32, device="cuda")
    K.rank_random(V, 11, a)
    K.rank_random(V, 11, b)
    K.rank_random(V, 12, c)
    torch.cuda.synchronize()
    assert np.array_equal(np.sort(host(a)), np.arange(V, dtype=np.uint32))
    assert np.array_equal(host(a), host(b))
    if V > 100:
        assert not np.array_equal(host(a), host(c))
        assert not np.array_equal(host(a), np.arange(V, dtype=np.uint32))


# ---------------------------------------------------------------------------------------------
# cache_by_heuristic ranking (toolkit/cache/cache_by_heuristic.cc:28-91)
# ---------------------------------------------------------------------------------------------
def run_heuristic(K, indptr, indices, train):
    V, E = len(indptr) - 1, len(indices)
    rank = torch.full((V,), -1, dtype=torch.int32, device="cuda")
    n_nbr = K.rank_by_heuristic(dev(indptr), dev(indices), V, E, dev(train), len(train), rank)
    torch.cuda.synchronize()
    assert n_nbr == int(np.diff(indptr.astype(np.int64))[train].sum())
    return host(rank)


def test_heuristic_ranking_matches_reference_tool_fixture(K, oracle):
    """Same dataset the reference's own cache-by-heuristic binary was run on (tests/golden/make_golden_tools.py)."""
    import os
    from fgnn_b200.synth import make_dataset_numpy
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tools_golden.npz"))
    ds = make_dataset_numpy(tuple(int(x) for x in g["spec"]), seed=int(g["seed"]))
    assert int(ds["indices"].astype(np.int64).sum()) == int(g["indices_sum"])
    got = run_heuristic(K, ds["indptr"], ds["indices"], ds["train_set"])
    assert np.array_equal(got, g["cache_by_heuristic"])
    # and the degree ranking of the same fixture
    V, E = len(ds["indptr"]) - 1, len(ds["indices"])
    deg = torch.empty(V, dtype=torch.int32, device="cuda")
    rank = torch.empty(V, dtype=torch.int32, device="cuda")
    ws = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device="cuda")
    K.rank_by_degree(dev(ds["indices"]), E, V, deg, rank, ws)
    torch.cuda.synchronize()
    assert np.array_equal(host(rank), g["cache_by_degree"])


@pytest.mark.parametrize("graph,n_train", [("small", 0), ("small", 1), ("small", 300), ("medium", 5000),
                                           ("medium", 1 << 16)])
def test_heuristic_ranking_matches_oracle(K, oracle, gs, gm, graph, n_train):
    g = gs if graph == "small" else gm
    V = len(g.indptr_np) - 1
    train = np.random.default_rng(n_train + 7).permutation(V)[:n_train].astype(np.uint32)
    got = run_heuristic(K, g.indptr_np, g.indices_np, train)
    exp = oracle.rank_by_heuristic(g.indptr_np, g.indices_np, train)
    assert np.array_equal(got, exp)
    assert np.array_equal(np.sort(got), np.arange(V, dtype=np.uint32))      # a permutation of the vertices
    assert np.array_equal(got[:n_train], train)                              # training nodes first, in order
