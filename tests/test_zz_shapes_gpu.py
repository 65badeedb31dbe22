"""BASELINE.json configs #3 and #4 at THEIR OWN dataset shapes, through the production calls:

  #3  PinSAGE random walks (3 layers, 4 walks of length 3, top-5, restart 0.5; train_pinsage.py:122-126) on the
      twitter-shaped graph: V = 41 652 230, E = 1 468 365 182, 256-d rows (datagen/twitter.sh:35-43)
  #4  GCN fanout [5,10,15] on the uk-2006-05-shaped graph: V = 77 741 046, E = 2 965 197 340, 256-d rows
      (datagen/uk-2006-05.sh:35-43) — `indptr` values exceed 2^31, so every index computation must be unsigned

The CPU oracle cannot finish these sizes in seconds: the checks are the size-independent properties of
tests/test_zz_fullsize_configs_gpu.py (every edge a CSR edge, exact per-seed edge counts, unique-list chain,
determinism) plus a bit-exact check of the extracted 1 KiB feature rows (hit + miss rows mixed).  One graph is
resident at a time (12 GB of topology for uk-2006-05)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
torch = pytest.importorskip("torch")

BATCH = 8000


def build(shape):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a B200-class HBM")
    from fgnn_b200 import kernels as K
    from fgnn_b200.synth import SHAPES, SEED, make_graph_torch
    K.load()
    V, E, D, C, T = SHAPES[shape]
    torch.cuda.empty_cache()
    indptr, indices = make_graph_torch(V, E, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + 1)
    train = torch.randperm(V, generator=g, device="cuda")[:T].to(torch.int32)
    rows = 1 << 20                                            # SAMGRAPH_EMPTY_FEAT-style 2^k-row host table
    gh = torch.Generator()
    gh.manual_seed(5)
    host_feat = (torch.rand((rows, D), generator=gh, dtype=torch.float32) * 2 - 1).pin_memory()
    return dict(V=V, E=E, D=D, T=T, indptr=indptr, indices=indices, train=train, host_feat=host_feat, mask=rows - 1)


@pytest.fixture(scope="module")
def twitter():
    g = build("twitter")
    yield g
    g.clear()
    torch.cuda.empty_cache()


@pytest.fixture(scope="module")
def uk():
    g = build("uk-2006-05")
    yield g
    g.clear()
    torch.cuda.empty_cache()


def check_extraction(big, hp, slot, cache_pct, freq_rank):
    """fgnn_k_gather_cached at D = 256: out[i] == host_table[n2o[i] & mask], bit for bit"""
    from test_zz_fullsize_configs_gpu import u64
    row_bytes = big["D"] * 4
    hp.build_cache(freq_rank, cache_pct, big["host_feat"], row_bytes, big["mask"])
    hp.stats.zero_()
    hp.gather(slot)
    torch.cuda.synchronize()
    sl = hp.slots[slot]
    n = int(sl.num_items.item())
    nodes = u64(sl.n2o[:n])
    got = hp.feat_out[:n].view(torch.float32).view(n, big["D"])
    for lo in range(0, n, 1 << 18):                           # the expected rows come from the pinned host table
        hi = min(n, lo + (1 << 18))
        exp = big["host_feat"][(nodes[lo:hi] & big["mask"]).cpu()].cuda()
        assert torch.equal(got[lo:hi].view(torch.int32), exp.view(torch.int32))
    hits, misses = [int(x) for x in hp.stats.tolist()]
    assert hits + misses == n
    return hits, misses


def test_twitter_shape_pinsage_random_walk(twitter):
    from fgnn_b200.pipeline import HotPath
    from test_zz_fullsize_configs_gpu import check_chain, check_random_walk_layer, run_batch, u64
    big = twitter
    assert (big["V"], big["E"], big["D"]) == (41652230, 1468365182, 256)
    rw = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4)
    Kn, L = 5, 3
    hp = HotPath(big["indptr"], big["indices"], big["V"], [Kn] * L, BATCH, "random_walk", seed=0xBEEF, rw=rw,
                 num_slots=1)
    assert hp.max_nodes == 1728000                                          # PredictNumNodes, SURVEY §8
    seeds = big["train"][:BATCH].contiguous()
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
    # extraction of the batch's 1 KiB rows with a 25 % cache ranked by in-batch order (any ranking works here)
    rank = torch.randperm(big["V"], device="cuda").to(torch.int32)
    hits, misses = check_extraction(big, hp, 0, 0.25, rank)
    assert hits > 0 and misses > 0
    sl, counts_again = run_batch(hp, seeds, 11)
    assert np.array_equal(counts, counts_again) and torch.equal(sl.n2o[:n_items], keep)


def test_uk_shape_gcn_three_layers_indptr_above_2_31(uk):
    from fgnn_b200.pipeline import HotPath
    from test_zz_fullsize_configs_gpu import check_chain, check_uniform_layer, run_batch, u64
    big = uk
    assert (big["V"], big["E"], big["D"]) == (77741046, 2965197340, 256)
    indptr64 = u64(big["indptr"])
    assert int(indptr64[-1]) == big["E"] and big["E"] > 2 ** 31
    assert bool((indptr64[1:] >= indptr64[:-1]).all())          # monotone as UNSIGNED 32-bit values
    fanouts = [5, 10, 15]
    hp = HotPath(big["indptr"], big["indices"], big["V"], fanouts, BATCH, "khop2", seed=0xFACE, num_slots=2)
    assert hp.max_nodes == 8448000
    # seeds: the train set's first batch plus the LAST vertices, whose rows start beyond 2^31
    seeds = big["train"][:BATCH].clone()
    tail = torch.arange(big["V"] - 500, big["V"], device="cuda", dtype=torch.int32)
    seeds[:500] = tail
    seeds = torch.unique(seeds)
    n_seed = seeds.numel()
    assert int((indptr64[u64(seeds)] >= 2 ** 31).sum()) >= 500
    hp.sample(seeds, n_seed, 3, slot=0)
    torch.cuda.synchronize()
    sl = hp.slots[0]
    counts = sl.counts.cpu().numpy().astype(np.int64)
    n_items = int(sl.num_items.item())
    n2o = u64(sl.n2o[:n_items])
    assert torch.equal(n2o[:n_seed], u64(seeds)) and torch.unique(n2o).numel() == n_items
    assert counts[2][0] == n_seed and counts[0][2] == n_items
    for i in (2, 1):
        assert counts[i][2] == counts[i - 1][0]
    high_rows = 0
    for i in (2, 1, 0):
        n_dst, n_edge, n_src = counts[i]
        row, col = u64(sl.row[i][:n_edge]), u64(sl.col[i][:n_edge])
        check_uniform_layer(big, n2o, row, col, n_dst, n_src, fanouts[i])
        high_rows += int((indptr64[n2o[:n_dst]] >= 2 ** 31).sum())
    assert high_rows > 1000                                     # sampled rows on both sides of the 2^31 boundary
    # the same batch inside a super-batch of two (slot 1) is bit-identical
    keep_n2o, keep_counts = sl.n2o[:n_items].clone(), counts.copy()
    other = big["train"][BATCH:2 * BATCH].contiguous()
    hp.sample_multi([(other, BATCH, 4, 0), (seeds, n_seed, 3, 1)])
    torch.cuda.synchronize()
    s1 = hp.slots[1]
    assert np.array_equal(s1.counts.cpu().numpy().astype(np.int64), keep_counts)
    assert torch.equal(s1.n2o[:n_items], keep_n2o)
    rank = torch.randperm(big["V"], device="cuda").to(torch.int32)
    hits, misses = check_extraction(big, hp, 1, 0.25, rank)
    assert hits > 0 and misses > 0
