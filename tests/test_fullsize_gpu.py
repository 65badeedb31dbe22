"""Full-size parity (BASELINE.json configs[1] shape): one B200, papers100M-shaped synthetic graph
(111 M vertices, 1.6 G edges, 128-d rows), mini-batches of 8000 seeds through the production calls
(fgnn_k_sample_batch + fgnn_k_gather_cached).  The CPU oracle cannot finish this size in seconds, so the checks are
the size-independent properties the domain offers; the checker on the device is plain torch indexing.

  * every TrainGraph edge (row, col) is a real CSR edge: n2o[row] in adj(n2o[col])
  * uniform without replacement: seed s contributes exactly min(deg(s), fanout) edges and no neighbour id more
    often than its row holds it
  * ordered unique: n2o starts with the seeds, holds no duplicate, and every layer's num_src/num_dst chain
  * determinism: the same (seed, batch_key) reproduces the batch bit for bit, another key does not
  * extraction: out[i] == table[n2o[i] & mask] bit for bit at 25 % cache (hit and miss rows mixed) and 100 %
  * PreSC: ranking is a permutation of [0, V) and freq[ranking] is non-increasing, ties by larger id first
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

FANOUTS = [25, 10]
BATCH = 8000


@pytest.fixture(scope="module")
def big():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs a B200-class HBM")
    from fgnn_b200 import kernels as K
    from fgnn_b200.synth import SHAPES, SEED, make_graph_torch
    K.load()
    V, E, D, C, T = SHAPES["papers100M"]
    indptr, indices = make_graph_torch(V, E, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(SEED + 1)
    train = torch.randperm(V, generator=g, device="cuda")[:T].to(torch.int32)
    rows = 1 << 20                                            # SAMGRAPH_EMPTY_FEAT-style 2^k-row host table
    gh = torch.Generator()
    gh.manual_seed(5)
    host_feat = (torch.rand((rows, D), generator=gh, dtype=torch.float32) * 2 - 1).pin_memory()
    return dict(V=V, E=E, D=D, T=T, indptr=indptr, indices=indices, train=train, host_feat=host_feat, mask=rows - 1)


def u64(t):
    """int32 tensor holding uint32 bits -> int64 values"""
    return t.to(torch.int64) & 0xFFFFFFFF


def sample(hp, seeds, key, slot=0):
    hp.sample(seeds, len(seeds), key, slot=slot)
    torch.cuda.synchronize()
    sl = hp.slots[slot]
    counts = sl.counts.cpu().numpy().astype(np.int64)         # [L][3] = num_dst, num_edge, num_src
    return sl, counts


@pytest.mark.parametrize("sample_type", ["khop2", "khop0"])
def test_fullsize_sample_batch_properties(big, sample_type):
    from fgnn_b200.pipeline import HotPath
    hp = HotPath(big["indptr"], big["indices"], big["V"], FANOUTS, BATCH, sample_type, seed=0xABCDEF, num_slots=2)
    seeds = big["train"][:BATCH].contiguous()
    sl, counts = sample(hp, seeds, 7)
    n_items = int(sl.num_items.item())
    n2o = u64(sl.n2o[:n_items])
    # ordered unique
    assert torch.equal(n2o[:BATCH], u64(seeds))
    assert torch.unique(n2o).numel() == n_items
    assert counts[1][0] == BATCH and counts[1][2] == counts[0][0] and counts[0][2] == n_items
    indptr = u64(big["indptr"])
    for i, f in ((1, FANOUTS[1]), (0, FANOUTS[0])):
        n_dst, n_edge, n_src = counts[i]
        row, col = u64(sl.row[i][:n_edge]), u64(sl.col[i][:n_edge])
        assert int(row.max()) < n_src and int(col.max()) < n_dst
        # seed-major compact order and exactly min(deg, f) edges per seed
        assert bool((col[1:] >= col[:-1]).all())
        gid = n2o[:n_dst]
        deg = indptr[gid + 1] - indptr[gid]
        assert torch.equal(torch.bincount(col, minlength=n_dst), torch.clamp(deg, max=f))
        # every edge is a CSR edge and, per (seed, neighbour id), is sampled at most as often as the row holds it
        # (positions are drawn without replacement; multi-edges may legitimately repeat an id)
        owner = torch.repeat_interleave(torch.arange(n_dst, device="cuda"), deg)
        first = torch.cumsum(deg, 0) - deg
        offs = torch.arange(owner.numel(), device="cuda") - first[owner]
        nbr = u64(big["indices"][indptr[gid][owner] + offs])
        au, ac = torch.unique(owner * (1 << 32) + nbr, return_counts=True)
        su, sc = torch.unique(col * (1 << 32) + n2o[row], return_counts=True)
        at = torch.searchsorted(au, su)
        assert bool((at < au.numel()).all())
        assert torch.equal(au[at], su)
        assert bool((sc <= ac[at]).all())
        del owner, first, offs, nbr, au, ac, su, sc, at
    # determinism / key sensitivity (slot 1 has its own table and scratch)
    keep = [sl.row[0][:counts[0][1]].clone(), sl.n2o[:n_items].clone()]
    sl1, counts1 = sample(hp, seeds, 7, slot=1)
    assert np.array_equal(counts, counts1)
    assert torch.equal(sl1.row[0][:counts[0][1]], keep[0]) and torch.equal(sl1.n2o[:n_items], keep[1])
    sl2, counts2 = sample(hp, seeds, 8, slot=1)
    assert not (np.array_equal(counts, counts2) and torch.equal(sl2.n2o[:n_items], keep[1]))


def test_fullsize_presc_cache_and_gather(big):
    from fgnn_b200 import kernels as K
    from fgnn_b200.pipeline import HotPath
    V, D = big["V"], big["D"]
    hp = HotPath(big["indptr"], big["indices"], V, FANOUTS, BATCH, "khop2", seed=0x1234, num_slots=1)
    freq = torch.zeros(V, dtype=torch.int32, device="cuda")
    for s in range(12):                                       # a slice of the pre-sampling epoch
        seeds = big["train"][s * BATCH:(s + 1) * BATCH].contiguous()
        hp.sample(seeds, BATCH, 1000 + s)
        hp.presample_count(freq)
    rank = torch.empty(V, dtype=torch.int32, device="cuda")
    ws = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device="cuda")
    K.presc_rank(freq, V, rank, ws)
    torch.cuda.synchronize()
    del ws
    r64 = u64(rank)
    fr = freq[r64].to(torch.int64)
    assert bool((fr[1:] <= fr[:-1]).all())
    tie = fr[1:] == fr[:-1]
    assert bool((r64[1:][tie] < r64[:-1][tie]).all())         # ties: larger id first (u64 {freq,id} descending)
    chk = torch.zeros(V, dtype=torch.bool, device="cuda")
    chk[r64] = True
    assert bool(chk.all())
    del chk, fr, tie
    seeds = big["train"][20 * BATCH:21 * BATCH].contiguous()
    dev_feat = big["host_feat"].cuda()
    for pct in (0.25, 1.0):
        hp.cache = hp.feat_out = None
        torch.cuda.empty_cache()
        hp.build_cache(rank, pct, big["host_feat"], D * 4, big["mask"])
        hp.stats.zero_()
        hp.sample(seeds, BATCH, 99)
        hp.gather(0)
        torch.cuda.synchronize()
        n = int(hp.slots[0].num_items.item())
        ids = u64(hp.slots[0].n2o[:n])
        got = hp.feat_out[:n].view(torch.int32)
        exp = dev_feat[ids & big["mask"]].view(torch.int32)   # CPUExtract semantics, checker = torch indexing
        assert torch.equal(got, exp)
        hits, misses = hp.stats.tolist()
        assert hits + misses == n
        in_cache = int((u64(hp.cache_table)[ids] != 0xFFFFFFFF).sum())
        assert hits == in_cache
        if pct == 1.0:
            assert misses == 0
