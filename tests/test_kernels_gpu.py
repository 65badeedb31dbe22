"""GPU parity tests: every kernel of the C-ABI (include/fgnn_kernels.h) against
the CPU oracle on the same seeded inputs.  Integer/byte work -> bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

SEED = 0x1234ABCD5678EF01


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fgnn_b200 import kernels
    kernels.load()  # raises if the extension is missing: no fallback
    return kernels


def dev(a):
    from fgnn_b200.synth import u32_tensor
    return u32_tensor(np.ascontiguousarray(a, dtype=np.uint32))


def host(t, n=None):
    from fgnn_b200.synth import to_np_u32
    a = to_np_u32(t)
    return a if n is None else a[:n]


def pick_seeds(indptr, n, seed, unique=True):
    rng = np.random.default_rng(seed)
    V = len(indptr) - 1
    if unique:
        return rng.permutation(V)[:n].astype(np.uint32)
    return rng.integers(0, V, size=n).astype(np.uint32)


class G:
    def __init__(self, indptr, indices):
        self.indptr_np, self.indices_np = indptr, indices
        self.indptr, self.indices = dev(indptr), dev(indices)


@pytest.fixture(scope="module")
def gs(graph_small):
    return G(*graph_small)


@pytest.fixture(scope="module")
def gm(graph_medium):
    return G(*graph_medium)


def run_khop(K, g, seeds, fanout, variant, batch_key=5, tag=1, d_n=None, n_max=None):
    n_max = len(seeds) if n_max is None else n_max
    inp = dev(np.concatenate([seeds, np.zeros(n_max - len(seeds), np.uint32)]))
    cap = max(1, n_max * fanout)
    out_src = torch.empty(cap, dtype=torch.int32, device="cuda")
    out_dst = torch.empty(cap, dtype=torch.int32, device="cuda")
    out_loc = torch.empty(cap, dtype=torch.int32, device="cuda")
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = K.new_chain_ws()
    dn = None
    if d_n is not None:
        dn = torch.tensor([d_n], dtype=torch.int32, device="cuda")
    K.sample_khop(variant, g.indptr, g.indices, inp, n_max, dn, fanout, K.rng(SEED, batch_key, tag),
                  out_src, out_dst, out_loc, num, ws)
    torch.cuda.synchronize()
    m = int(num.item())
    assert int(ws.abs().sum().item()) == 0, "chain workspace must be left zeroed"
    return host(out_src, m), host(out_dst, m), host(out_loc, m)


@pytest.mark.parametrize("variant,fanout", [(2, 5), (2, 10), (2, 25), (2, 32), (2, 40), (2, 100),
                                            (0, 5), (0, 15), (0, 25)])
@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 2500])
def test_sample_khop_matches_oracle(K, oracle, gs, variant, fanout, n):
    seeds = pick_seeds(gs.indptr_np, n, 3)
    s, d, l = run_khop(K, gs, seeds, fanout, variant)
    fn = oracle.sample_khop2 if variant == 2 else oracle.sample_khop0
    es, ed = fn(gs.indptr_np, gs.indices_np, seeds, fanout, SEED, 5, 1)
    assert np.array_equal(s, es)
    assert np.array_equal(d, ed)
    # local src = index of the seed in the input list
    if len(s):
        assert np.array_equal(seeds[l], s)
        assert np.all(np.diff(l.astype(np.int64)) >= 0)


@pytest.mark.parametrize("variant", [0, 2])
def test_sample_khop_device_count_and_large(K, oracle, gm, variant):
    seeds = pick_seeds(gm.indptr_np, 40000, 9)
    # device-side count smaller than the bound
    s, d, _ = run_khop(K, gm, seeds[:33333], 10, variant, batch_key=77, tag=0, d_n=33333, n_max=40000)
    fn = oracle.sample_khop2 if variant == 2 else oracle.sample_khop0
    es, ed = fn(gm.indptr_np, gm.indices_np, seeds[:33333], 10, SEED, 77, 0)
    assert np.array_equal(s, es) and np.array_equal(d, ed)


def test_sample_khop_hub_rows(K, oracle):
    """Rows far above the CTA-cooperative threshold (reservoir) and above fanout (F-Y)."""
    rng = np.random.default_rng(5)
    degs = np.array([0, 1, 70000, 3, 25, 26, 9000, 5000, 4121, 4122], np.uint32)
    indptr = np.concatenate([[0], np.cumsum(degs)]).astype(np.uint32)
    indices = rng.integers(0, len(degs), size=int(indptr[-1])).astype(np.uint32)
    g = G(indptr, indices)
    seeds = np.arange(len(degs), dtype=np.uint32)
    for variant, fn in ((0, oracle.sample_khop0), (2, oracle.sample_khop2)):
        s, d, _ = run_khop(K, g, seeds, 25, variant)
        es, ed = fn(indptr, indices, seeds, 25, SEED, 5, 1)
        assert np.array_equal(s, es) and np.array_equal(d, ed)


def test_sample_khop2_without_replacement_property(K, gm):
    """Size-independent property: positions drawn for a row are distinct -> on a graph whose
    rows hold distinct ids, each sampled row has no duplicate and is a subset of the row."""
    V = 5000
    rng = np.random.default_rng(1)
    degs = rng.integers(0, 200, size=V).astype(np.uint32)
    indptr = np.concatenate([[0], np.cumsum(degs)]).astype(np.uint32)
    indices = np.concatenate([rng.permutation(100000)[:d] for d in degs]).astype(np.uint32)
    g = G(indptr, indices)
    seeds = np.arange(V, dtype=np.uint32)
    for variant in (0, 2):
        s, d, l = run_khop(K, g, seeds, 15, variant)
        counts = np.bincount(l, minlength=V)
        assert np.array_equal(counts, np.minimum(degs, 15))
        key = l.astype(np.uint64) << np.uint64(32) | d.astype(np.uint64)
        assert len(np.unique(key)) == len(key)
        # subset check
        row_key = np.repeat(np.arange(V, dtype=np.uint64), degs) << np.uint64(32) | indices.astype(np.uint64)
        assert np.all(np.isin(key, row_key))


# ---------------------------------------------------------------------------
class HT:
    def __init__(self, K, max_items):
        self.K = K
        self.cap = K.ht_capacity(max_items)
        self.table = torch.empty(K.ht_bytes(self.cap) // 4, dtype=torch.int32, device="cuda")
        self.n2o = torch.empty(max_items + 1, dtype=torch.int32, device="cuda")
        self.num = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.ws = K.new_chain_ws()
        K.ht_reset(self.table, self.cap, self.num)

    def reset(self):
        self.K.ht_reset(self.table, self.cap, self.num)

    def fill_unique(self, ids):
        self.K.ht_fill_unique(self.table, self.cap, dev(ids), len(ids), None, self.n2o, self.num)

    def fill_duplicates(self, ids):
        d = dev(ids)
        pos = torch.empty(max(1, len(ids)), dtype=torch.int32, device="cuda")
        self.K.ht_fill_duplicates(self.table, self.cap, d, len(ids), None, pos, self.n2o, self.num, self.ws)
        return d, pos

    def num_items(self):
        torch.cuda.synchronize()
        return int(self.num.item())

    def unique(self):
        return host(self.n2o, self.num_items())

    def map(self, d=None, pos=None, n=None):
        out = torch.empty(max(1, n), dtype=torch.int32, device="cuda")
        self.K.ht_map(self.table, self.cap, d, pos, n, None, out)
        torch.cuda.synchronize()
        return host(out, n)


@pytest.mark.parametrize("n_seed,n_dup,universe", [(0, 0, 10), (8, 0, 100), (100, 5000, 300), (8000, 200000, 50000),
                                                   (1000, 300000, 1 << 20), (1, 100000, 3)])
def test_hashtable_matches_oracle(K, oracle, n_seed, n_dup, universe):
    rng = np.random.default_rng(n_seed + n_dup)
    seeds = rng.permutation(universe)[:n_seed].astype(np.uint32)
    max_items = n_seed + 3 * n_dup + 16
    ht = HT(K, max_items)
    oh = oracle.hashtable(max_items)
    ht.fill_unique(seeds)
    oh.fill_unique(seeds)
    for rnd in range(3):  # three "layers"
        ids = rng.integers(0, universe, size=n_dup).astype(np.uint32)
        d, pos = ht.fill_duplicates(ids)
        oh.fill_duplicates(ids)
        assert ht.num_items() == oh.num_items
        assert np.array_equal(ht.unique(), oh.unique())
        if n_dup:
            exp = oh.map(ids)
            assert np.array_equal(ht.map(None, pos, n_dup), exp)   # via remembered bucket
            assert np.array_equal(ht.map(d, None, n_dup), exp)     # via probe
        assert int(ht.ws.abs().sum().item()) == 0
    # reset -> empty again, local ids restart at 0
    ht.reset()
    assert ht.num_items() == 0
    ht.fill_unique(seeds)
    assert np.array_equal(ht.unique(), seeds)


@pytest.mark.parametrize("n_seed,n_dup,universe", [(0, 1, 10), (8, 33, 100), (100, 5000, 300), (8000, 200000, 50000),
                                                   (1000, 1500000, 1 << 20), (1, 100000, 3)])
def test_hashtable_fill_and_map_fused_matches_oracle(K, oracle, n_seed, n_dup, universe):
    """fgnn_k_ht_fill_duplicates_map == FillWithDuplicates followed by GPUMapEdges: same unique order, and
    every item of the fill remapped in the same pass (owners in lower chunks are waited for, not re-probed);
    1.5 M items exercise chunks longer than the register-cached tiles."""
    rng = np.random.default_rng(n_seed * 3 + n_dup)
    seeds = rng.permutation(universe)[:n_seed].astype(np.uint32)
    max_items = n_seed + 3 * n_dup + 16
    ht = HT(K, max_items)
    oh = oracle.hashtable(max_items)
    ht.fill_unique(seeds)
    oh.fill_unique(seeds)
    for rnd in range(3):
        ids = rng.integers(0, universe, size=n_dup).astype(np.uint32)
        d = dev(ids)
        pos = torch.empty(n_dup, dtype=torch.int32, device="cuda")
        loc = torch.full((n_dup + 1,), -1, dtype=torch.int32, device="cuda")
        m = max(0, n_dup - rnd)                                                 # ragged device count
        dn = torch.tensor([m], dtype=torch.int32, device="cuda")
        K.ht_fill_duplicates_map(ht.table, ht.cap, d, n_dup, dn, pos, ht.n2o, ht.num, loc, ht.ws)
        oh.fill_duplicates(ids[:m])
        assert ht.num_items() == oh.num_items
        assert np.array_equal(ht.unique(), oh.unique())
        assert np.array_equal(host(loc, m), oh.map(ids[:m]))
        assert host(loc)[m:].tolist() == [0xFFFFFFFF] * (n_dup + 1 - m)
        assert int(ht.ws.abs().sum().item()) == 0


def test_hashtable_map_absent_is_empty(K):
    ht = HT(K, 100)
    ht.fill_unique(np.array([5, 6, 7], np.uint32))
    q = dev(np.array([7, 8, 5], np.uint32))
    out = ht.map(q, None, 3)
    assert out.tolist() == [2, 0xFFFFFFFF, 0]


# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n,V,pct", [(0, 100, 0.5), (1, 100, 0.0), (1000, 5000, 0.25), (100000, 70000, 0.1),
                                     (5000, 5000, 1.0)])
def test_cache_table_and_split(K, oracle, n, V, pct):
    rng = np.random.default_rng(n + V)
    rank = rng.permutation(V).astype(np.uint32)
    nc = oracle.num_cached(V, pct)
    table_ref = oracle.cache_table_build(rank, V, nc)
    table = torch.empty(V, dtype=torch.int32, device="cuda")
    K.cache_table_build(table, V, dev(rank), nc)
    torch.cuda.synchronize()
    assert np.array_equal(host(table), table_ref)
    nodes = rng.integers(0, V, size=n).astype(np.uint32)
    bufs = [torch.empty(max(1, n), dtype=torch.int32, device="cuda") for _ in range(4)]
    counts = torch.zeros(2, dtype=torch.int32, device="cuda")
    ws = K.new_chain_ws()
    K.cache_split(table, dev(nodes), n, None, bufs[0], bufs[1], bufs[2], bufs[3], counts, ws)
    torch.cuda.synchronize()
    nm, nh = counts.tolist()
    ms, md, cs, cd = oracle.cache_split(table_ref, nodes)
    assert (nm, nh) == (len(ms), len(cs)) and nm + nh == n   # cuda_loops.cc:999 invariant
    assert np.array_equal(host(bufs[0], nm), ms) and np.array_equal(host(bufs[1], nm), md)
    assert np.array_equal(host(bufs[2], nh), cs) and np.array_equal(host(bufs[3], nh), cd)
    assert int(ws.abs().sum().item()) == 0


# ---------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,dim", [(np.float32, 128), (np.float32, 100), (np.float32, 256), (np.float32, 1),
                                       (np.int64, 1), (np.uint8, 7), (np.float64, 3), (np.int16, 5), (np.float32, 602)])
@pytest.mark.parametrize("n", [0, 1, 1000, 33333])
def test_row_copy_extract(K, oracle, dtype, dim, n):
    rng = np.random.default_rng(dim + n)
    V = 5000
    if np.issubdtype(dtype, np.floating):
        src = rng.standard_normal((V, dim)).astype(dtype)
    else:
        src = rng.integers(0, 100, size=(V, dim)).astype(dtype)
    idx = rng.integers(0, V, size=n).astype(np.uint32)
    exp = oracle.extract(src, idx)
    d_src = torch.from_numpy(src.view(np.uint8).reshape(V, -1)).cuda()
    row_bytes = d_src.shape[1]
    out = torch.zeros((max(1, n), row_bytes), dtype=torch.uint8, device="cuda")
    K.row_copy(out, None, d_src, dev(idx), n, None, row_bytes)
    torch.cuda.synchronize()
    got = out[:n].cpu().numpy().reshape(n, row_bytes).view(dtype).reshape(n, dim)
    assert np.array_equal(got.view(np.uint8), exp.view(np.uint8))


def test_row_copy_scatter_and_mask(K, oracle):
    rng = np.random.default_rng(0)
    V, n, dim = 1 << 10, 3000, 64
    src = rng.standard_normal((V, dim)).astype(np.float32)
    idx = rng.integers(0, 1 << 20, size=n).astype(np.uint32)       # masked like SAMGRAPH_EMPTY_FEAT=10
    dst_idx = rng.permutation(n).astype(np.uint32)
    exp = np.zeros((n, dim), np.float32)
    oracle.row_copy(exp, dst_idx, src, idx, n, dim * 4, mask=(1 << 10) - 1)
    out = torch.zeros((n, dim), dtype=torch.float32, device="cuda")
    K.row_copy(out, dev(dst_idx), torch.from_numpy(src).cuda(), dev(idx), n, None, dim * 4, src_mask=(1 << 10) - 1)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), exp.view(np.uint32))


@pytest.mark.parametrize("num_shards", [1, 2, 4])
@pytest.mark.parametrize("dim,pct", [(128, 0.3), (100, 0.0), (256, 1.0)])
def test_gather_cached_matches_reference_pipeline(K, oracle, num_shards, dim, pct):
    """Fused gather == GetMissCacheIndex + ExtractMissData + CombineMiss + CombineCache."""
    rng = np.random.default_rng(dim)
    V, n = 20000, 50000
    feat = rng.standard_normal((V, dim)).astype(np.float32)
    rank = rng.permutation(V).astype(np.uint32)
    nc = oracle.num_cached(V, pct)
    table_ref = oracle.cache_table_build(rank, V, nc)
    nodes = rng.integers(0, V, size=n).astype(np.uint32)
    # reference pipeline on the CPU oracle
    ms, md, cs, cd = oracle.cache_split(table_ref, nodes)
    cache = oracle.extract(feat, rank[:nc])                    # dist_cache_manager_host.cc:98-109
    exp = np.zeros((n, dim), np.float32)
    miss_rows = oracle.extract(feat, ms)
    oracle.row_copy(exp, md, miss_rows, None, len(ms), dim * 4)
    oracle.row_copy(exp, cd, cache, cs, len(cs), dim * 4)
    assert np.array_equal(exp, feat[nodes])
    # device: cache striped over `num_shards` buffers (owner = slot % T, row = slot // T)
    shards = []
    for t in range(num_shards):
        rows = cache[t::num_shards]
        shards.append(torch.from_numpy(np.ascontiguousarray(rows)).cuda() if len(rows) else
                      torch.zeros((1, dim), dtype=torch.float32, device="cuda"))
    ptrs = torch.tensor([s.data_ptr() for s in shards], dtype=torch.int64, device="cuda")
    table = dev(table_ref)
    host_feat = torch.from_numpy(feat).pin_memory()             # misses read over UVA
    out = torch.zeros((n, dim), dtype=torch.float32, device="cuda")
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    K.gather_cached(out, dev(nodes), n, None, table, ptrs, num_shards, host_feat, dim * 4, stats)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), exp.view(np.uint32))
    assert stats.tolist() == [len(cs), len(ms)]


@pytest.mark.parametrize("impl", ["bulk", "group", "flat"])
@pytest.mark.parametrize("row_bytes,n,n_max,pct,shards", [
    (512, 0, 64, 0.5, 1), (512, 1, 1, 0.5, 1), (512, 3, 3, 1.0, 1), (512, 33, 40, 0.0, 1),
    (512, 70001, 70001, 0.9, 1), (512, 70001, 90000, 1.0, 3), (400, 12345, 12345, 0.7, 2),
    (16, 5000, 5000, 0.5, 1), (48, 5000, 6000, 0.5, 1), (1024, 9999, 9999, 0.8, 1), (2048, 3001, 3001, 0.6, 2),
    (4096, 2000, 2000, 0.5, 1), (6400, 777, 777, 0.5, 1), (12288, 301, 301, 0.5, 1), (36, 2000, 2000, 0.5, 1)])
def test_gather_cached_impls_edge_cases(K, oracle, impl, row_bytes, n, n_max, pct, shards):
    """Every implementation of fgnn_k_gather_cached (TMA bulk ring, warp-group, flat) over ragged counts
    (device count < n_max, n not a multiple of the row group), rows from 16 B to 12 KB (ring-depth
    fallbacks), non-16-byte rows (flat path), striped shards and host-resident misses: bit-exact."""
    import os
    rng = np.random.default_rng(row_bytes * 7 + n)
    V = 30000
    src = rng.integers(0, 256, size=(V, row_bytes), dtype=np.uint8)
    rank = rng.permutation(V).astype(np.uint32)
    nc = oracle.num_cached(V, pct)
    table_ref = oracle.cache_table_build(rank, V, nc)
    nodes = rng.integers(0, V, size=n_max).astype(np.uint32)
    exp = src[nodes[:n]]
    cache = src[rank[:nc]]
    bufs = []
    for t in range(shards):
        rows = cache[t::shards]
        bufs.append(torch.from_numpy(np.ascontiguousarray(rows)).cuda() if len(rows) else
                    torch.zeros((1, row_bytes), dtype=torch.uint8, device="cuda"))
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    host_src = torch.from_numpy(src).pin_memory()
    out = torch.full((n_max + 1, row_bytes), 0xA5, dtype=torch.uint8, device="cuda")
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    d_n = torch.tensor([n], dtype=torch.int32, device="cuda")
    old = {k: os.environ.get(k) for k in ("FGNN_TUNING_DYNAMIC", "FGNN_GATHER_IMPL")}
    os.environ["FGNN_TUNING_DYNAMIC"] = "1"
    os.environ["FGNN_GATHER_IMPL"] = impl
    try:
        K.gather_cached(out, dev(nodes), n_max, d_n, dev(table_ref), ptrs, shards, host_src, row_bytes, stats)
        torch.cuda.synchronize()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    got = out.cpu().numpy()
    assert np.array_equal(got[:n], exp)
    assert (got[n:] == 0xA5).all(), "rows beyond the device count must not be written"
    hit = int((table_ref[nodes[:n]] != 0xFFFFFFFF).sum())
    assert stats.tolist() == [hit, n - hit]


@pytest.mark.parametrize("impl", ["bulk", "bulk+defer", "group", "flat"])
@pytest.mark.parametrize("row_bytes,shards,self_shard,pct,repl", [(512, 4, 1, 0.8, 0.2), (512, 2, 0, 1.0, 0.5),
                                                                 (1024, 8, 7, 0.6, 0.05), (400, 3, 2, 0.7, 0.3),
                                                                 (512, 1, 0, 0.5, 0.5), (512, 4, 3, 0.3, 0.6),
                                                                 (4096, 4, 0, 0.9, 0.0)])
def test_gather_cached_hybrid_layout(K, oracle, impl, row_bytes, shards, self_shard, pct, repl, monkeypatch):
    """fgnn_k_gather_cached_layout: the hottest R slots in a local replica, slots >= R striped over `shards`
    buffers ((s-R) % T, (s-R) // T), misses from pinned host memory: bit-exact rows; the kernel's remote-row
    counter equals the number of gathered rows whose stripe owner is not `self_shard`."""
    from fgnn_b200 import partition as P
    monkeypatch.setenv("FGNN_TUNING_DYNAMIC", "1")
    defer = impl == "bulk+defer"            # peer rows listed by the ring kernel and copied by a second launch
    monkeypatch.setenv("FGNN_GATHER_IMPL", "bulk" if defer else impl)
    monkeypatch.setenv("FGNN_GATHER_DEFER", "1" if defer else "0")
    rng = np.random.default_rng(row_bytes + shards)
    V, n = 30000, 40001
    src = rng.integers(0, 256, size=(V, row_bytes), dtype=np.uint8)
    rank = rng.permutation(V).astype(np.uint32)
    nc = oracle.num_cached(V, pct)
    R = min(nc, int(V * repl))
    table_ref = oracle.cache_table_build(rank, V, nc)
    nodes = rng.integers(0, V, size=n).astype(np.uint32)
    cache = src[rank[:nc]]
    replica = torch.from_numpy(np.ascontiguousarray(cache[:R])).cuda() if R else None
    bufs = []
    for t in range(shards):
        rows = cache[R + t::shards]
        assert len(rows) == P.stripe_rows(nc - R, shards, t)
        bufs.append(torch.from_numpy(np.ascontiguousarray(rows)).cuda() if len(rows) else
                    torch.zeros((1, row_bytes), dtype=torch.uint8, device="cuda"))
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    host_src = torch.from_numpy(src).pin_memory()
    d_table, d_nodes = dev(table_ref), dev(nodes)
    lay = K.CacheLayout()
    lay.table, lay.shards, lay.num_shards, lay.self_shard = d_table.data_ptr(), ptrs.data_ptr(), shards, self_shard
    lay.replica, lay.num_replicated = (replica.data_ptr() if R else None), R
    lay.miss_src, lay.miss_mask, lay.row_bytes = host_src.data_ptr(), 0xFFFFFFFFFFFFFFFF, row_bytes
    out = torch.zeros((n, row_bytes), dtype=torch.uint8, device="cuda")
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    remote = torch.zeros(1, dtype=torch.int64, device="cuda")
    dws = None
    if defer:
        dws = torch.zeros(int(K.load().fgnn_k_gather_defer_workspace_bytes(n)), dtype=torch.uint8, device="cuda")
        lay.defer_ws = dws.data_ptr()
    for rep in range(2 if defer else 1):     # twice: the list counter must be re-armed by the second pass
        out.zero_()
        stats.zero_()
        remote.zero_()
        K.gather_cached_layout(out, d_nodes, n, None, lay, stats, remote)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), src[nodes])
    if defer:
        assert int(dws[:8].sum().item()) == 0
    slots = table_ref[nodes]
    hit = slots != 0xFFFFFFFF
    assert stats.tolist() == [int(hit.sum()), int((~hit).sum())]
    striped = hit & (slots >= R)
    owners = (slots[striped].astype(np.int64) - R) % shards
    assert int(remote.item()) == (int((owners != self_shard).sum()) if shards > 1 else 0)
    for s_ in (0, R - 1, R, R + 1, nc - 1):
        if 0 <= s_ < nc:
            o, r = P.slot_owner(int(s_), shards, R)
            assert (o is None and r == s_) if s_ < R else (o == (s_ - R) % shards and r == (s_ - R) // shards)


# ---------------------------------------------------------------------------
def test_presc_count_and_rank(K, oracle):
    rng = np.random.default_rng(3)
    V = 100000
    freq_ref = np.zeros(V, np.uint32)
    freq = torch.zeros(V, dtype=torch.int32, device="cuda")
    for _ in range(4):
        nodes = (rng.zipf(1.3, size=60000) % V).astype(np.uint32)
        oracle.freq_count(freq_ref, nodes)
        K.freq_count(freq, dev(nodes), len(nodes), None)
    torch.cuda.synchronize()
    assert np.array_equal(host(freq), freq_ref)
    ws = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device="cuda")
    rank = torch.empty(V, dtype=torch.int32, device="cuda")
    K.presc_rank(freq, V, rank, ws)
    torch.cuda.synchronize()
    assert np.array_equal(host(rank), oracle.presc_rank(freq_ref))


# ---------------------------------------------------------------------------
def weight_tables(oracle, indptr, indices, seed=2):
    rng = np.random.default_rng(seed)
    w = rng.integers(1, 11, size=len(indices)).astype(np.float32)
    prob, alias = oracle.build_alias_table(indptr, indices, w)
    prefix = oracle.build_prefix_table(indptr, w)
    return prob, alias, prefix


@pytest.mark.parametrize("kind", [1, 2, 4])
@pytest.mark.parametrize("ranked", [True, False])
@pytest.mark.parametrize("n,fanout,unique", [(0, 5, True), (1, 10, True), (700, 10, True), (2500, 25, True),
                                             (500, 7, False)])
def test_sample_replace_matches_oracle(K, oracle, gs, kind, n, fanout, unique, ranked):
    """ranked = the seeds are ordered by the rank-by-bitmap kernels (default in the engine), else by the CUB sort;
    the rank workspace must come back all-zero.  Duplicate seeds (unique=False) are outside the bitmap's contract
    (a layer's inputs are unique): they only run on the sort path."""
    if ranked and not unique:
        pytest.skip("rank-by-bitmap needs unique seeds (always true for a layer's inputs)")
    prob, alias, prefix = weight_tables(oracle, gs.indptr_np, gs.indices_np)
    seeds = pick_seeds(gs.indptr_np, n, 4, unique)
    n_max = n + 13
    inp = dev(np.concatenate([seeds, np.zeros(n_max - n, np.uint32)]))
    cap = n_max * fanout
    outs = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(3)]
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    dn = torch.tensor([n], dtype=torch.int32, device="cuda")
    wsb = torch.empty(K.sample_replace_workspace_bytes(n_max, fanout), dtype=torch.uint8, device="cuda")
    ws = K.new_chain_ws()
    fprob, falias, fprefix = (torch.from_numpy(prob).cuda(), dev(alias), torch.from_numpy(prefix).cuda())
    V = len(gs.indptr_np) - 1
    rank_ws = K.new_rank_ws(V) if ranked else None
    for rep in range(2 if ranked else 1):      # twice: the first call must leave the rank workspace clean
        K.sample_replace(kind, gs.indptr, gs.indices, fprob, falias, fprefix, inp, n_max, dn, fanout,
                         K.rng(SEED, 9, 2), outs[0], outs[1], outs[2], num, wsb, ws, rank_ws=rank_ws, num_nodes=V)
        torch.cuda.synchronize()
    if ranked:
        words = (V + 31) // 32
        nbk = (V + 1023) // 1024
        al = lambda x: (x + 255) & ~255
        assert int(rank_ws[:al(words * 4) + al((nbk + 1) * 4) + 256].to(torch.int64).sum().item()) == 0
    m = int(num.item())
    if kind == 1:
        es, ed = oracle.sample_khop1(gs.indptr_np, gs.indices_np, seeds, fanout, SEED, 9, 2)
    elif kind == 2:
        es, ed = oracle.sample_weighted_khop(gs.indptr_np, gs.indices_np, prob, alias, seeds, fanout, SEED, 9, 2)
    else:
        es, ed = oracle.sample_weighted_khop_prefix(gs.indptr_np, gs.indices_np, prefix, seeds, fanout, SEED, 9, 2)
    assert m == len(es)
    assert np.array_equal(host(outs[0], m), es) and np.array_equal(host(outs[1], m), ed)
    if m:
        assert np.array_equal(seeds[host(outs[2], m)], es)
    assert int(ws.abs().sum().item()) == 0


@pytest.mark.parametrize("n,fanout", [(0, 5), (3, 5), (900, 10), (2500, 25)])
def test_sample_weighted_hash_dedup_matches_oracle(K, oracle, gs, n, fanout):
    prob, alias, _ = weight_tables(oracle, gs.indptr_np, gs.indices_np)
    seeds = pick_seeds(gs.indptr_np, n, 6)
    cap = max(1, n * fanout)
    outs = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(3)]
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    ws = K.new_chain_ws()
    K.sample_weighted_hash_dedup(gs.indptr, gs.indices, torch.from_numpy(prob).cuda(), dev(alias), dev(seeds), n,
                                 None, fanout, K.rng(SEED, 3, 0), outs[0], outs[1], outs[2], num, ws)
    torch.cuda.synchronize()
    m = int(num.item())
    es, ed = oracle.sample_weighted_khop(gs.indptr_np, gs.indices_np, prob, alias, seeds, fanout, SEED, 3, 0,
                                         hash_dedup=True)
    assert m == len(es)
    assert np.array_equal(host(outs[0], m), es) and np.array_equal(host(outs[1], m), ed)


@pytest.mark.parametrize("n,W,L,Kn,p", [(0, 4, 3, 5, 0.5), (1, 4, 3, 5, 0.5), (3000, 4, 3, 5, 0.5), (777, 3, 3, 4, 0.3),
                                        (500, 6, 8, 5, 0.1), (300, 10, 10, 7, 0.05), (40, 40, 30, 10, 0.02),
                                        (500, 8, 4, 10, 0.0), (500, 1, 1, 1, 0.9), (300, 5, 2, 3, 1.0)])
def test_random_walk_topk_matches_oracle(K, oracle, gs, n, W, L, Kn, p):
    seeds = pick_seeds(gs.indptr_np, n, 8)
    cap = max(1, n * max(Kn, W * L))
    outs = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(4)]
    tmps = [torch.empty(cap, dtype=torch.int32, device="cuda") for _ in range(2)]
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    wsb = torch.empty(K.sample_random_walk_workspace_bytes(max(1, n), Kn), dtype=torch.uint8, device="cuda")
    ws = K.new_chain_ws()
    K.sample_random_walk(gs.indptr, gs.indices, dev(seeds), n, None, L, p, W, Kn, K.rng(SEED, 21, 2), outs[0],
                         outs[1], outs[2], outs[3], num, tmps[0], tmps[1], wsb, ws)
    torch.cuda.synchronize()
    m = int(num.item())
    ts, td = oracle.random_walk(gs.indptr_np, gs.indices_np, seeds, L, p, W, SEED, 21, 2)
    assert np.array_equal(host(tmps[0], n * W * L), ts)
    live = ts != 0xFFFFFFFF
    assert np.array_equal(host(tmps[1], n * W * L)[live], td[live])
    es, ed, ew = oracle.topk(ts, td, seeds, W * L, Kn)
    assert m == len(es)
    assert np.array_equal(host(outs[0], m), es) and np.array_equal(host(outs[1], m), ed)
    assert np.array_equal(host(outs[3], m), ew)
    assert int(ws.abs().sum().item()) == 0


@pytest.mark.parametrize("sample_type,fanouts,n_seed", [
    ("khop2", [5, 10, 15], 2000), ("khop2", [25, 10], 777), ("khop0", [5, 10], 1500), ("khop1", [10, 5], 900),
    ("weighted_khop", [10, 5], 900), ("weighted_khop_prefix", [8, 4], 500), ("weighted_khop_hash_dedup", [6, 3], 400),
    ("random_walk", [5, 5, 5], 300), ("khop2", [3], 0), ("khop2", [4, 4], 1)])
@pytest.mark.parametrize("fuse,grid_div", [(14, 1), (2, 1), (6, 1), (0, 1), (6, 16), (2, 16)])
def test_sample_batch_call_matches_oracle_driver(K, oracle, gs, gm, sample_type, fanouts, n_seed, fuse, grid_div,
                                                 monkeypatch):
    """fgnn_k_sample_batch (the one C call the engine makes per mini-batch: fused sample+insert and
    compact+remap, counts written by the kernels) against the numpy restatement of DoGPUSample
    (cuda_loops.cc:50-267), for every SampleType; two slots used alternately on two streams."""
    from fgnn_b200.pipeline import HotPath
    from oracle.oracle import sample_batch_oracle
    if (fuse, grid_div) != (14, 1) and sample_type != "khop2":
        pytest.skip("FGNN_BATCH_FUSE only changes the uniform k-hop kernel sequence")
    # FGNN_GRID_DIV = 16 shrinks every persistent grid: chunks of the chained scans span many tiles
    monkeypatch.setenv("FGNN_TUNING_DYNAMIC", "1")
    monkeypatch.setenv("FGNN_GRID_DIV", str(grid_div))
    # bit 1: remap folded into the compaction pass; bit 2: padded sampler + one dual-count chained scan;
    # bit 3: the two-launch-per-layer chain of fast_chain.cu (default)
    monkeypatch.setenv("FGNN_BATCH_FUSE", str(fuse))
    g = gm if sample_type in ("khop2", "khop0") else gs
    graph = dict(indptr=g.indptr_np, indices=g.indices_np)
    kw = {}
    rw = None
    if sample_type.startswith("weighted"):
        prob, alias, prefix = weight_tables(oracle, g.indptr_np, g.indices_np)
        graph.update(prob_table=prob, alias_table=alias, prob_prefix_table=prefix)
        kw = dict(prob_table=torch.from_numpy(prob).cuda(), alias_table=dev(alias),
                  prefix_table=torch.from_numpy(prefix).cuda())
    if sample_type == "random_walk":
        rw = dict(random_walk_length=3, random_walk_restart_prob=0.5, num_random_walk=4, num_neighbor=fanouts[0])
    batch = max(n_seed, 8)
    hp = HotPath(g.indptr, g.indices, len(g.indptr_np) - 1, fanouts, batch, sample_type, seed=SEED, rw=rw,
                 num_slots=2, **kw)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    runs = []
    for rep in range(4):                     # slots reused: scratch and chain workspace must come back clean
        seeds = pick_seeds(g.indptr_np, n_seed, 100 + rep)
        slot = rep % 2
        d_seeds = dev(seeds) if n_seed else torch.zeros(1, dtype=torch.int32, device="cuda")
        torch.cuda.current_stream().synchronize()
        with torch.cuda.stream(streams[slot]):
            hp.sample(d_seeds, n_seed, 40 + rep, slot=slot)
        runs.append((seeds, slot, 40 + rep, d_seeds))
        if rep % 2 == 1:                     # two batches in flight, then check both
            torch.cuda.synchronize()
            for seeds_, slot_, key_, _ in runs[-2:]:
                exp = sample_batch_oracle(oracle, graph, seeds_, fanouts, sample_type, SEED, key_, rw=rw)
                sl = hp.slots[slot_]
                cnt = sl.counts.cpu().numpy().astype(np.int64)
                n_in = int(sl.num_items.item())
                assert np.array_equal(host(sl.n2o, n_in), exp["input_nodes"])
                for i in range(len(fanouts)):
                    lay = exp["layers"][i]
                    assert cnt[i].tolist() == [lay["num_dst"], lay["num_edge"], lay["num_src"]]
                    m = lay["num_edge"]
                    assert np.array_equal(host(sl.row[i], m), lay["row"])
                    assert np.array_equal(host(sl.col[i], m), lay["col"])
                    if sample_type == "random_walk":
                        assert np.array_equal(host(sl.data[i], m), lay["data"])
                assert int(sl.chain.abs().sum().item()) == 0


def check_slot(oracle, graph, hp, slot, seeds, fanouts, key):
    from oracle.oracle import sample_batch_oracle
    exp = sample_batch_oracle(oracle, graph, seeds, fanouts, "khop2", SEED, key)
    sl = hp.slots[slot]
    cnt = sl.counts.cpu().numpy().astype(np.int64)
    n_in = int(sl.num_items.item())
    assert np.array_equal(host(sl.n2o, n_in), exp["input_nodes"])
    for i in range(len(fanouts)):
        lay = exp["layers"][i]
        assert cnt[i].tolist() == [lay["num_dst"], lay["num_edge"], lay["num_src"]]
        m = lay["num_edge"]
        assert np.array_equal(host(sl.row[i], m), lay["row"])
        assert np.array_equal(host(sl.col[i], m), lay["col"])
    assert int(sl.chain.abs().sum().item()) == 0


@pytest.mark.parametrize("graph_kind", ["medium", "dense", "hub"])
@pytest.mark.parametrize("fanouts,n_seed,k_super", [([25, 10], 777, 4), ([5, 10, 15], 1500, 3), ([3], 0, 2),
                                                    ([4, 4], 1, 1), ([25, 10], 3000, 8), ([40, 33], 300, 2)])
@pytest.mark.parametrize("versioned", [1, 0])
def test_sample_batch_multi_matches_oracle(K, oracle, gm, graph_kind, fanouts, n_seed, k_super, versioned, monkeypatch):
    """fgnn_k_sample_batch_multi: k_super mini-batches (own table / scratch / outputs each) in ONE call, every
    layer two launches for all of them; versioned table reset (no memset between batches) and the cleared-table
    mode; batches of different sizes in one call; slots reused across calls.  Bit-exact vs the oracle driver.
    dense = most rows longer than the fanout (several Fisher-Yates rounds per tile); hub = a few 20k-neighbour
    rows among isolated and short ones."""
    from fgnn_b200.pipeline import HotPath
    monkeypatch.setenv("FGNN_HT_VERSIONED", str(versioned))
    monkeypatch.delenv("FGNN_BATCH_FUSE", raising=False)
    if graph_kind == "medium":
        g = gm
    elif graph_kind == "dense":
        from conftest import small_graph
        g = G(*small_graph(4000, 400000, seed=21, zero_deg_frac=0.01))
    else:
        rng = np.random.default_rng(5)
        V = 30000
        deg = rng.integers(0, 6, size=V)
        deg[rng.permutation(V)[:12]] = 20000
        indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
        indices = rng.integers(0, V, size=int(indptr[-1])).astype(np.uint32)
        g = G(indptr, indices)
    graph = dict(indptr=g.indptr_np, indices=g.indices_np)
    batch = max(n_seed, 8)
    hp = HotPath(g.indptr, g.indices, len(g.indptr_np) - 1, fanouts, batch, "khop2", seed=SEED, num_slots=k_super)
    for rep in range(3):                      # slots reused: versions advance, workspaces must come back clean
        batches, keep = [], []
        for k in range(k_super):
            n_k = n_seed if k % 2 == 0 else n_seed // 2          # ragged super-batch
            seeds = pick_seeds(g.indptr_np, n_k, 1000 + 10 * rep + k)
            if graph_kind == "hub" and n_k >= 12:
                hubs = np.nonzero(np.diff(g.indptr_np.astype(np.int64)) == 20000)[0].astype(np.uint32)
                seeds = np.concatenate([hubs, np.setdiff1d(seeds, hubs)])[:n_k].astype(np.uint32)
            d_seeds = dev(seeds) if n_k else torch.zeros(1, dtype=torch.int32, device="cuda")
            batches.append((d_seeds, n_k, 500 + 10 * rep + k, k))
            keep.append(seeds)
        hp.sample_multi(batches)
        torch.cuda.synchronize()
        for k in range(k_super):
            check_slot(oracle, graph, hp, k, keep[k], fanouts, 500 + 10 * rep + k)


def test_versioned_table_wraps(K, oracle, gs):
    """130 consecutive mini-batches on one slot: the 7-bit version tag wraps once (the table is cleared by
    fgnn_k_ht_next_version), every batch bit-exact; stale buckets of earlier batches must read as free."""
    from fgnn_b200.pipeline import HotPath
    graph = dict(indptr=gs.indptr_np, indices=gs.indices_np)
    fanouts = [6, 4]
    hp = HotPath(gs.indptr, gs.indices, len(gs.indptr_np) - 1, fanouts, 64, "khop2", seed=SEED, num_slots=1)
    assert hp.versioned
    seen = set()
    for rep in range(130):
        seeds = pick_seeds(gs.indptr_np, 64, 3000 + rep)
        hp.sample(dev(seeds), 64, 9000 + rep)
        seen.add(int(hp.slots[0].plan.version))
        if rep % 13 == 0 or rep > 124:
            torch.cuda.synchronize()
            check_slot(oracle, graph, hp, 0, seeds, fanouts, 9000 + rep)
    assert min(seen) == 1 and max(seen) == 126


# ---------------------------------------------------------------------------
def test_full_batch_pipeline_matches_oracle_driver(K, oracle, gm):
    """sample -> unique -> remap for 3 layers, all counts device-resident, one sync at the end;
    compared with the numpy restatement of DoGPUSample (cuda_loops.cc:50-267)."""
    from oracle.oracle import sample_batch_oracle
    fanouts = [5, 10, 15]
    seeds = pick_seeds(gm.indptr_np, 2000, 12)
    exp = sample_batch_oracle(oracle, dict(indptr=gm.indptr_np, indices=gm.indices_np), seeds, fanouts, "khop2",
                              SEED, 42)
    bound = oracle.predict_num_nodes(len(seeds), fanouts)
    ht = HT(K, bound)
    ht.fill_unique(seeds)
    ws = K.new_chain_ws()
    cur, cur_n_dev, cur_max = ht.n2o, ht.num, len(seeds)
    layers = [None] * 3
    for i in (2, 1, 0):
        f = fanouts[i]
        cap = cur_max * f
        dst = torch.empty(cap, dtype=torch.int32, device="cuda")
        col = torch.empty(cap, dtype=torch.int32, device="cuda")
        row = torch.empty(cap, dtype=torch.int32, device="cuda")
        pos = torch.empty(cap, dtype=torch.int32, device="cuda")
        ne = torch.zeros(1, dtype=torch.int32, device="cuda")
        # layer input = the first *cur_n_dev entries of the running unique list
        n_in = cur_n_dev.clone()
        K.sample_khop(2, gm.indptr, gm.indices, cur, cur_max, n_in, f, K.rng(SEED, 42, i), None, dst, col, ne, ws)
        K.ht_fill_duplicates(ht.table, ht.cap, dst, cap, ne, pos, ht.n2o, ht.num, ws)
        K.ht_map(ht.table, ht.cap, None, pos, cap, ne, row)
        layers[i] = (row, col, ne, n_in, ht.num.clone())
        cur_max = min(bound, cur_max * (f + 1))
    torch.cuda.synchronize()
    for i in range(3):
        row, col, ne, n_in, n_src = layers[i]
        m = int(ne.item())
        e = exp["layers"][i]
        assert m == e["num_edge"] and int(n_in.item()) == e["num_dst"] and int(n_src.item()) == e["num_src"]
        assert np.array_equal(host(row, m), e["row"]) and np.array_equal(host(col, m), e["col"])
    assert np.array_equal(ht.unique(), exp["input_nodes"])


@pytest.mark.parametrize("n", [0, 1, 1000, 300001])
def test_epoch_shuffle_matches_oracle(K, oracle, n):
    train = np.random.default_rng(n).permutation(1 << 20)[:n].astype(np.uint32)
    out = torch.empty(max(1, n), dtype=torch.int32, device="cuda")
    ws = torch.empty(K.shuffle_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    for epoch in (0, 3):
        K.shuffle(dev(train), n, SEED, epoch, out, ws)
        torch.cuda.synchronize()
        got = host(out, n)
        assert np.array_equal(got, oracle.shuffle(train, SEED, epoch))
        assert np.array_equal(np.sort(got), np.sort(train))          # a permutation
    if n > 1000:
        assert not np.array_equal(oracle.shuffle(train, SEED, 0), oracle.shuffle(train, SEED, 3))


# ---------------------------------------------------------------------------------------------
# CSC hand-off: indptr / indices / edge ids of one block == stable counting sort by dst
# ---------------------------------------------------------------------------------------------
def _csc_expected(row, col, num_dst):
    perm = np.argsort(col, kind="stable").astype(np.uint32)
    indptr = np.searchsorted(col[perm], np.arange(num_dst + 1), side="left").astype(np.uint32)
    return indptr, row[perm], perm


@pytest.mark.parametrize("e,num_dst,num_src,sorted_col,live", [
    (0, 0, 0, True, None), (0, 7, 5, False, None), (1, 1, 1, True, None), (1, 4, 9, False, None),
    (1000, 64, 500, True, None), (1000, 64, 500, False, None), (5000, 5000, 777, False, 3100),
    (5000, 300, 777, True, 1), (70000, 8000, 60000, True, 65537), (70000, 8000, 60000, False, None),
    (1 << 20, 88000, 550000, False, (1 << 20) - 12345), (1 << 20, 88000, 550000, True, None)])
def test_coo_to_csc_matches_oracle(K, oracle, e, num_dst, num_src, sorted_col, live):
    rng = np.random.default_rng(e * 31 + num_dst)
    n_live = e if live is None else live
    col = rng.integers(0, max(num_dst, 1), size=e).astype(np.uint32)      # some dst nodes get no edge at all
    if num_dst > 3:
        col[col == num_dst - 1] = 0                                        # the last dst is always empty
    if sorted_col:
        col[:n_live] = np.sort(col[:n_live])
    row = rng.integers(0, max(num_src, 1), size=e).astype(np.uint32)
    d_row, d_col = dev(row), dev(col)
    d_e = None if live is None else dev(np.array([n_live], np.uint32))
    indptr = torch.full((num_dst + 1,), -1, dtype=torch.int32, device="cuda")
    indices = torch.full((max(e, 1),), -1, dtype=torch.int32, device="cuda")
    eids = torch.full((max(e, 1),), -1, dtype=torch.int32, device="cuda")
    K.coo_to_csc(d_row, d_col, e, d_e, num_dst, sorted_col, indptr, indices, eids)
    torch.cuda.synchronize()
    exp_indptr, exp_indices, exp_eids = _csc_expected(row[:n_live], col[:n_live], num_dst)
    if n_live <= 5000:   # the oracle's plain counting sort (python loops) pins the vectorised expectation
        o_indptr, o_indices, o_eids = oracle.coo_to_csc(row[:n_live], col[:n_live], num_dst)
        assert np.array_equal(o_indptr, exp_indptr) and np.array_equal(o_indices, exp_indices)
        assert np.array_equal(o_eids, exp_eids)
    assert np.array_equal(host(indptr), exp_indptr)
    assert np.array_equal(host(indices, n_live), exp_indices)
    assert np.array_equal(host(eids, n_live), exp_eids)
    if sorted_col:
        assert np.array_equal(exp_eids, np.arange(n_live, dtype=np.uint32))
        # zero-copy form used by the runtime: no indices / edge ids asked for, no workspace
        indptr2 = torch.full((num_dst + 1,), -1, dtype=torch.int32, device="cuda")
        K.coo_to_csc(d_row, d_col, e, d_e, num_dst, True, indptr2)
        torch.cuda.synchronize()
        assert np.array_equal(host(indptr2), exp_indptr)
    if e > n_live:        # nothing is written past the live count
        assert (host(indices)[n_live:e] == 0xFFFFFFFF).all() and (host(eids)[n_live:e] == 0xFFFFFFFF).all()
