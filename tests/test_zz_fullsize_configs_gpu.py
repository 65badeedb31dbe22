"""Full-size properties for the other BASELINE.json configurations, on the papers100M-shaped graph (111 M vertices,
1.6 G edges) through the production call fgnn_k_sample_batch:
  #4-like  GCN fanout [5,10,15] (three layers, bound 8.45 M nodes per batch), khop2
  #3-like  PinSAGE random walks: 3 layers x top-5 of 4 walks of length 3, restart 0.5
  #5-like  weighted k-hop (alias tables built on the GPU from kDefault weights 1..10), GraphSAGE [25,10]
The oracle cannot finish these sizes in seconds; the checks are size-independent properties, the checker on the
device is plain torch.  (File name sorts last on purpose: these ran for the first time in the round-end GPU suite.)"""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(420)]   # pytest-timeout: a hang must not eat the GPU run
torch = pytest.importorskip("torch")

BATCH = 8000


@pytest.fixture(scope="module")
def big():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a B200-class HBM")
    from fgnn_b200 import kernels as K
    from fgnn_b200.synth import SHAPES, SEED, make_graph_torch
    K.load()
    V, E, D, C, T = SHAPES["papers100M"]
    indptr, indices = make_graph_torch(V, E, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + 1)
    train = torch.randperm(V, generator=g, device="cuda")[:T].to(torch.int32)
    return dict(V=V, E=E, indptr=indptr, indices=indices, train=train)


def u64(t):
    return t.to(torch.int64) & 0xFFFFFFFF


def run_batch(hp, seeds, key, slot=0):
    hp.sample(seeds, len(seeds), key, slot=slot)
    torch.cuda.synchronize()
    sl = hp.slots[slot]
    return sl, sl.counts.cpu().numpy().astype(np.int64)      # [L][3] = num_dst, num_edge, num_src


def check_chain(counts, n_items, L):
    """num_src / num_dst chain of DoGPUSample (cuda_loops.cc:87-229): layer L-1 starts from the seeds, every layer's
    inputs are the previous layer's unique list, the last unique list is input_nodes."""
    assert counts[L - 1][0] == BATCH and counts[0][2] == n_items
    for i in range(L - 1, 0, -1):
        assert counts[i][2] == counts[i - 1][0]
        assert counts[i][2] >= counts[i][0]


def csr_pair_keys(big, gid):
    """sorted unique keys owner_index * 2^32 + neighbour id, with multiplicities, for the rows of vertices `gid`"""
    indptr = u64(big["indptr"])
    dev = gid.device
    deg = indptr[gid + 1] - indptr[gid]
    owner = torch.repeat_interleave(torch.arange(gid.numel(), device=dev), deg)
    first = torch.cumsum(deg, 0) - deg
    offs = torch.arange(owner.numel(), device=dev) - first[owner]
    nbr = u64(big["indices"][indptr[gid][owner] + offs])
    return deg, torch.unique(owner * (1 << 32) + nbr, return_counts=True)


# ---- per-layer property checkers (device-agnostic: tests/test_fullsize_checkers_cpu.py runs them on oracle batches) ----
def check_uniform_layer(big, n2o, row, col, n_dst, n_src, fanout):
    """uniform without replacement, seed-major: exactly min(deg, f) edges per seed, every edge a CSR edge, no
    neighbour id more often than its row holds it."""
    assert int(row.max()) < n_src and int(col.max()) < n_dst
    assert bool((col[1:] >= col[:-1]).all())
    deg, (au, ac) = csr_pair_keys(big, n2o[:n_dst])
    assert torch.equal(torch.bincount(col, minlength=n_dst), torch.clamp(deg, max=fanout))
    su, sc = torch.unique(col * (1 << 32) + n2o[row], return_counts=True)
    at = torch.searchsorted(au, su)
    assert bool((at < au.numel()).all()) and torch.equal(au[at], su) and bool((sc <= ac[at]).all())


def check_random_walk_layer(big, n2o, row, col, data, n_dst, n_src, top_k, budget):
    """start-node major, at most K outputs per node, visit counts >= 1 that sum to at most W * L per node and do not
    increase inside a node, no id twice per node, isolated start nodes produce nothing."""
    assert int(row.max()) < n_src and int(col.max()) < n_dst
    assert bool((col[1:] >= col[:-1]).all())
    per_seed = torch.bincount(col, minlength=n_dst)
    assert int(per_seed.max()) <= top_k
    assert int(data.min()) >= 1
    visits = torch.zeros(n_dst, dtype=torch.int64, device=col.device).index_add_(0, col, data)
    assert int(visits.max()) <= budget
    same = col[1:] == col[:-1]
    assert bool((data[1:][same] <= data[:-1][same]).all())
    assert torch.unique(col * (1 << 32) + n2o[row]).numel() == row.numel()
    indptr = u64(big["indptr"])
    gid = n2o[:n_dst]
    assert int(per_seed[(indptr[gid + 1] - indptr[gid]) == 0].sum()) == 0


def check_weighted_layer(big, n2o, row, col, n_dst, n_src, fanout):
    """with replacement + sort by the seed's GLOBAL id + adjacent dedup: at most f edges per seed, at least one for
    every non-isolated seed, equal neighbours never adjacent, every edge a CSR edge."""
    assert int(row.max()) < n_src and int(col.max()) < n_dst
    src_gid = n2o[col]
    assert bool((src_gid[1:] >= src_gid[:-1]).all())
    same = col[1:] == col[:-1]
    assert bool((row[1:][same] != row[:-1][same]).all())
    per_seed = torch.bincount(col, minlength=n_dst)
    deg, (au, ac) = csr_pair_keys(big, n2o[:n_dst])
    assert int(per_seed.max()) <= fanout
    assert bool(((per_seed > 0) == (deg > 0)).all())
    su = torch.unique(col * (1 << 32) + n2o[row])
    # reference quirk kept bit for bit: an entry whose probability ended at exactly 1 has alias 0 (never written by
    # create_alias_table.cc) and `r < prob` fails when curand_uniform returns 1.0 (2^-25 per draw), so vertex 0 may
    # appear as a neighbour without being one (cuda_sampling_weighted_khop.cu:65-70)
    su = su[(su & 0xFFFFFFFF) != 0]
    at = torch.searchsorted(au, su).clamp(max=au.numel() - 1)
    assert torch.equal(au[at], su)


def test_fullsize_gcn_three_layers(big):
    from fgnn_b200.pipeline import HotPath
    fanouts = [5, 10, 15]
    hp = HotPath(big["indptr"], big["indices"], big["V"], fanouts, BATCH, "khop2", seed=0xFACE, num_slots=1)
    assert hp.max_nodes == 8448000                                          # PredictNumNodes, SURVEY §8
    seeds = big["train"][:BATCH].contiguous()
    sl, counts = run_batch(hp, seeds, 3)
    n_items = int(sl.num_items.item())
    n2o = u64(sl.n2o[:n_items])
    assert torch.equal(n2o[:BATCH], u64(seeds))
    assert torch.unique(n2o).numel() == n_items
    check_chain(counts, n_items, 3)
    for i in (2, 1, 0):
        n_dst, n_edge, n_src = counts[i]
        row, col = u64(sl.row[i][:n_edge]), u64(sl.col[i][:n_edge])
        check_uniform_layer(big, n2o, row, col, n_dst, n_src, fanouts[i])
    keep = sl.n2o[:n_items].clone()
    sl, counts_again = run_batch(hp, seeds, 3)
    assert np.array_equal(counts, counts_again) and torch.equal(sl.n2o[:n_items], keep)


def test_fullsize_pinsage_random_walk(big):
    from fgnn_b200.pipeline import HotPath
    rw = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4)
    Kn, L = 5, 3
    hp = HotPath(big["indptr"], big["indices"], big["V"], [Kn] * L, BATCH, "random_walk", seed=0xBEEF, rw=rw,
                 num_slots=1)
    seeds = big["train"][BATCH:2 * BATCH].contiguous()
    sl, counts = run_batch(hp, seeds, 11)
    n_items = int(sl.num_items.item())
    n2o = u64(sl.n2o[:n_items])
    assert torch.equal(n2o[:BATCH], u64(seeds)) and torch.unique(n2o).numel() == n_items
    check_chain(counts, n_items, L)
    budget = rw["num_random_walk"] * rw["random_walk_length"]
    for i in range(L - 1, -1, -1):
        n_dst, n_edge, n_src = counts[i]
        assert 0 < n_edge <= n_dst * Kn
        row, col, data = u64(sl.row[i][:n_edge]), u64(sl.col[i][:n_edge]), u64(sl.data[i][:n_edge])
        check_random_walk_layer(big, n2o, row, col, data, n_dst, n_src, Kn, budget)
    keep = sl.n2o[:n_items].clone()
    sl, counts_again = run_batch(hp, seeds, 11)
    assert np.array_equal(counts, counts_again) and torch.equal(sl.n2o[:n_items], keep)


def test_fullsize_weighted_khop(big):
    from fgnn_b200 import kernels as K
    from fgnn_b200.pipeline import HotPath
    V, E = big["V"], big["E"]
    fanouts = [25, 10]
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    weights = torch.randint(1, 11, (E,), generator=g, device="cuda", dtype=torch.int32).to(torch.float32)
    prob = torch.empty(E, dtype=torch.float32, device="cuda")
    alias = torch.empty(E, dtype=torch.int32, device="cuda")
    K.build_alias_table(big["indptr"], big["indices"], V, E, weights, prob, alias)
    torch.cuda.synchronize()
    del weights
    torch.cuda.empty_cache()
    # table sanity at full size: probabilities in (0, 1]; a zero alias only where the probability is exactly 1
    # or where node 0 really is the alias
    assert float(prob.min()) > 0.0 and float(prob.max()) <= 1.0
    hp = HotPath(big["indptr"], big["indices"], V, fanouts, BATCH, "weighted_khop", seed=0xD00D, prob_table=prob,
                 alias_table=alias, num_slots=1)
    seeds = big["train"][2 * BATCH:3 * BATCH].contiguous()
    sl, counts = run_batch(hp, seeds, 5)
    n_items = int(sl.num_items.item())
    n2o = u64(sl.n2o[:n_items])
    assert torch.equal(n2o[:BATCH], u64(seeds)) and torch.unique(n2o).numel() == n_items
    check_chain(counts, n_items, 2)
    for i in (1, 0):
        n_dst, n_edge, n_src = counts[i]
        row, col = u64(sl.row[i][:n_edge]), u64(sl.col[i][:n_edge])
        check_weighted_layer(big, n2o, row, col, n_dst, n_src, fanouts[i])
    keep = sl.n2o[:n_items].clone()
    sl, counts_again = run_batch(hp, seeds, 5)
    assert np.array_equal(counts, counts_again) and torch.equal(sl.n2o[:n_items], keep)
