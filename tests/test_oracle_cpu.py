"""CPU tests (no GPU): the oracle against published known-answer vectors, against the fixtures generated
from the reference's own CPU code (tests/golden/ref_cpu_golden.npz, made by tests/golden/make_golden.py),
against the live reference build (oracle/_ref) when present, and against hand-computed small cases."""
import os

import numpy as np
import pytest
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
E = 0xFFFFFFFF


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_cpu_golden.npz"))


# ---- RNG -----------------------------------------------------------------------------------------
def test_philox4x32_10_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert oracle.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle.philox([E] * 4, [E] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_rand_stream_layout(oracle):
    seed, bk = 0x0123456789ABCDEF, 0x00000007_00000042
    key = [seed & E, (seed >> 32) ^ (bk >> 32)]
    for item, draw, tag in [(0, 0, 0), (5, 3, 1), (77, 9, 2), (123456, 100, 0xFFFF0001)]:
        blk = oracle.philox([draw >> 2, item, tag, bk & E], key)
        assert oracle.rand_u32(seed, bk, tag, item, draw) == blk[draw & 3]


def test_uniform_formulas_match_curand_headers(oracle):
    L = oracle.lib
    assert L.fgo_uniform_f32(0) == np.float32(2.0 ** -33)              # curand_uniform.h:69-72
    assert L.fgo_uniform_f32(E) == np.float32(1.0)
    assert L.fgo_uniform_f32(1 << 31) == np.float32(0.5)
    assert L.fgo_uniform_f64(0, 0) == 2.0 ** -54                        # curand_uniform.h:101-106
    x, y = 0x12345678, 0x9ABCDEF0
    z = x ^ (y << 21)
    assert L.fgo_uniform_f64(x, y) == z * 2.0 ** -53 + 2.0 ** -54


# ---- sizing --------------------------------------------------------------------------------------
def test_predict_num_nodes_and_table_size(oracle, golden):
    assert oracle.predict_num_nodes(8000, [25, 10]) == int(golden["predict_8000_25_10"]) == 2288000
    assert oracle.predict_num_nodes(8000, [5, 10, 15]) == int(golden["predict_8000_5_10_15"]) == 8448000
    # TableSize = (1 << (1 + floor(log2(n >> 1)))) << scale   (cuda_hashtable.cu:125-128)
    assert oracle.table_size(2288000, 2) == (1 << 21) << 2
    assert oracle.table_size(8448000, 2) == (1 << 23) << 2
    assert oracle.table_size(12, 3) == 64                               # frequency hashmap: 12 walk steps


# ---- unique / remap vs the reference's hashtables ----------------------------------------------------
def test_hashtable_matches_reference_fixtures(oracle, golden):
    ht = oracle.hashtable(40000)
    ht.fill_unique(golden["ht_seeds"])
    for r in range(3):
        ids = golden["ht_round_%d" % r]
        uniq = ht.fill_duplicates(ids)
        for kind in (0, 2):   # CPUHashTable0 and single-threaded CPUHashTable2 agree on first-occurrence order
            assert np.array_equal(uniq, golden["ht%d_unique_%d" % (kind, r)])
            assert np.array_equal(ht.map(ids), golden["ht%d_map_src_%d" % (kind, r)])
            assert np.array_equal(ht.map(ids[::-1].copy()), golden["ht%d_map_dst_%d" % (kind, r)])


def test_hashtable_edge_cases(oracle):
    ht = oracle.hashtable(16)
    assert ht.num_items == 0
    ht.fill_duplicates(np.array([], np.uint32))
    assert ht.num_items == 0
    ht.fill_unique(np.array([9, 4], np.uint32))
    assert ht.fill_duplicates(np.array([4, 4, 7, 9, 7, 0], np.uint32)).tolist() == [9, 4, 7, 0]
    assert ht.map(np.array([0, 7, 4, 9], np.uint32)).tolist() == [3, 2, 1, 0]
    ht.reset()
    assert ht.num_items == 0 and ht.fill_duplicates(np.array([7], np.uint32)).tolist() == [7]


# ---- extraction ------------------------------------------------------------------------------------
def test_extract_matches_reference_fixtures(oracle, golden):
    assert np.array_equal(oracle.extract(golden["ex_feat"], golden["ex_index"]), golden["ex_feat_out"])
    assert np.array_equal(oracle.extract(golden["ex_label"], golden["ex_index"]), golden["ex_label_out"])


def test_row_copy_scatter_gather_mask(oracle):
    src = np.arange(40, dtype=np.float32).reshape(8, 5)
    dst = np.zeros((4, 5), np.float32)
    oracle.row_copy(dst, np.array([3, 0, 2], np.uint32), src, np.array([9, 2, 15], np.uint32), 3, 20, mask=7)
    assert np.array_equal(dst[3], src[1]) and np.array_equal(dst[0], src[2]) and np.array_equal(dst[2], src[7])
    assert not dst[1].any()


# ---- uniform samplers -----------------------------------------------------------------------------
@pytest.mark.parametrize("fanout", [5, 15])
def test_khop_layout_and_copy_path_match_reference_fixtures(oracle, golden, fanout):
    indptr, indices, inp = golden["g_indptr"], golden["g_indices"], golden["khop_input"]
    for fn in (oracle.sample_khop0, oracle.sample_khop2):
        s, d = fn(indptr, indices, inp, fanout, 1, 2, 3)
        assert np.array_equal(s, golden["khop_src_f%d" % fanout])        # seed-major compact COO
        small = golden["khop_small_mask_f%d" % fanout]
        assert np.array_equal(d[small], golden["khop_small_dst_f%d" % fanout])   # deg <= fanout: whole row


def test_khop_without_replacement_and_subset(oracle):
    rng = np.random.default_rng(0)
    degs = rng.integers(0, 60, size=500)
    indptr = np.concatenate([[0], np.cumsum(degs)]).astype(np.uint32)
    indices = np.concatenate([rng.permutation(5000)[:d] for d in degs]).astype(np.uint32)  # distinct per row
    seeds = np.arange(500, dtype=np.uint32)
    for fn in (oracle.sample_khop0, oracle.sample_khop2):
        s, d = fn(indptr, indices, seeds, 10, 9, 9, 0)
        assert np.array_equal(np.bincount(s, minlength=500), np.minimum(degs, 10))
        for v in (3, 77, 400):
            got = d[s == v]
            assert len(set(got.tolist())) == len(got)
            assert set(got.tolist()) <= set(indices[indptr[v]:indptr[v + 1]].tolist())


@pytest.mark.parametrize("which", ["khop0", "khop2"])
def test_uniform_sampler_distribution_matches_reference(oracle, golden, which):
    """The reference seeds its RNG itself (parity unpinned at the bit level): compare the per-neighbour
    pick frequencies of one hub row, chi-square against uniform-without-replacement and two-sample against
    the counts recorded from the reference's CPUSampleKHop0/2."""
    indptr, indices = golden["g_indptr"], golden["g_indices"]
    hub, reps = int(golden["dist_hub"]), int(golden["dist_reps"])
    fn = oracle.sample_khop0 if which == "khop0" else oracle.sample_khop2
    one = np.full(reps, hub, np.uint32)
    s, d = fn(indptr, indices, one, 10, 4242, 1, 0)
    assert len(d) == reps * 10
    row = indices[indptr[hub]:indptr[hub + 1]]
    ids, mult = np.unique(row, return_counts=True)
    got = np.zeros(len(ids))
    gi, gc = np.unique(d, return_counts=True)
    got[np.searchsorted(ids, gi)] = gc
    expected = mult / mult.sum() * len(d)
    keep = expected >= 5
    chi, p = stats.chisquare(got[keep], expected[keep] * got[keep].sum() / expected[keep].sum())
    assert p > 1e-4, "oracle %s deviates from uniform: p=%g" % (which, p)
    ref_cnt = np.zeros(len(ids))
    ref_cnt[np.searchsorted(ids, golden["dist_%s_ids" % which])] = golden["dist_%s_cnt" % which]
    table = np.stack([got[keep], ref_cnt[keep]])
    _, p2, _, _ = stats.chi2_contingency(table)
    assert p2 > 1e-4, "oracle %s vs reference frequencies: p=%g" % (which, p2)


def test_khop1_sorted_by_src_and_adjacent_dedup(oracle):
    indptr = np.array([0, 3, 3, 4], np.uint32)
    indices = np.array([10, 11, 12, 20], np.uint32)
    seeds = np.array([2, 1, 0], np.uint32)           # row 1 is empty, row 2 has a single neighbour
    s, d = oracle.sample_khop1(indptr, indices, seeds, 4, 5, 5, 0)
    # rows come out ordered by seed id; the 4 identical draws of row 2 collapse to one entry
    assert s.tolist()[-1] == 2 and d.tolist()[-1] == 20 and (s == 2).sum() == 1
    assert (s == 1).sum() == 0 and np.all(np.diff(s.astype(np.int64)) >= 0)
    row0 = d[s == 0]
    assert np.all(row0[1:] != row0[:-1]) and set(row0.tolist()) <= {10, 11, 12}


def test_weighted_sampler_follows_alias_table(oracle):
    rng = np.random.default_rng(1)
    indptr = np.array([0, 6], np.uint32)
    indices = np.array([5, 6, 7, 8, 9, 10], np.uint32)
    w = np.array([1, 2, 3, 4, 5, 10], np.float32)
    prob, alias = oracle.build_alias_table(indptr, indices, w)
    prefix = oracle.build_prefix_table(indptr, w)
    assert np.allclose(prefix, np.cumsum(w))
    # alias table invariant: sum_k [ prob_k * 1{id_k = v} + (1 - prob_k) * 1{alias_k = v} ] / len = w_v / sum w
    mass = np.zeros(11)
    for k in range(6):
        mass[indices[k]] += prob[k]
        mass[alias[k]] += 1 - prob[k]
    assert np.allclose(mass[5:] / 6, w / w.sum(), atol=1e-6)
    n = 20000
    for name in ("alias", "prefix"):
        cnt = np.zeros(6)
        for rep in range(n // 500):
            seeds = np.zeros(1, np.uint32)
            if name == "alias":
                # with-replacement draws; adjacent duplicates are dropped, so count via khop over many keys
                s, d = oracle.sample_weighted_khop(indptr, indices, prob, alias, seeds, 1, 7, rep, 0)
            else:
                s, d = oracle.sample_weighted_khop_prefix(indptr, indices, prefix, seeds, 1, 7, rep, 0)
            cnt[d[0] - 5] += 1
        # fanout 1 -> one draw per batch key; 40 keys is few, so only sanity-check support here
        assert cnt.sum() == n // 500
    # distribution over many independent draws: use fanout f with distinct items (one seed per draw)
    many = np.zeros(3000, np.uint32)
    s, d = oracle.sample_weighted_khop(indptr, indices, prob, alias, many, 1, 11, 3, 0)
    # duplicates of the same seed id are adjacent after the sort -> equal consecutive picks collapse; use the
    # hash-dedup-free primitive instead: draw through distinct single-row graphs
    big_indptr = (np.arange(3001) * 6).astype(np.uint32)
    big_indices = np.tile(indices, 3000)
    bprob, balias = np.tile(prob, 3000), np.tile(alias, 3000)
    s, d = oracle.sample_weighted_khop(big_indptr, big_indices, bprob, balias, np.arange(3000, dtype=np.uint32), 1, 11, 3, 0)
    cnt = np.bincount(d - 5, minlength=6)
    _, p = stats.chisquare(cnt, w / w.sum() * cnt.sum())
    assert p > 1e-4
    bpre = np.tile(prefix, 3000)
    s, d = oracle.sample_weighted_khop_prefix(big_indptr, big_indices, bpre, np.arange(3000, dtype=np.uint32), 1, 11, 3, 0)
    cnt = np.bincount(d - 5, minlength=6)
    _, p = stats.chisquare(cnt, w / w.sum() * cnt.sum())
    assert p > 1e-4


def test_weighted_hash_dedup_distinct(oracle):
    indptr = np.array([0, 8, 10], np.uint32)
    indices = np.array([1, 2, 3, 4, 5, 6, 7, 8, 30, 31], np.uint32)
    w = np.ones(10, np.float32)
    prob, alias = oracle.build_alias_table(indptr, indices, w)
    s, d = oracle.sample_weighted_khop(indptr, indices, prob, alias, np.array([0, 1], np.uint32), 5, 3, 3, 0, hash_dedup=True)
    assert (s == 0).sum() == 5 and len(set(d[s == 0].tolist())) == 5
    assert d[s == 1].tolist() == [30, 31]            # deg <= fanout: the whole row, in order


# ---- random walk + top-K ------------------------------------------------------------------------------
def test_topk_hand_case(oracle):
    # node 0: visits [7,8,7,9,8,7] -> counts 7:3, 8:2, 9:1 ; node 1: all dead ; node 2: ties keep first occurrence
    ts = np.array([100] * 6 + [E] * 6 + [300] * 6, np.uint32)
    td = np.array([7, 8, 7, 9, 8, 7] + [0] * 6 + [5, 6, 6, 5, 4, 4], np.uint32)
    s, d, c = oracle.topk(ts, td, np.array([100, 200, 300], np.uint32), 6, 2)
    assert s.tolist() == [100, 100, 300, 300]
    assert d.tolist() == [7, 8, 5, 6]                # 5,6,4 all have count 2: first occurrence order, keep 2
    assert c.tolist() == [3, 2, 2, 2]


def test_random_walk_semantics(oracle):
    # a ring 0->1->2->0 ... (each node has exactly one in-neighbour): walks are deterministic
    indptr = np.arange(5, dtype=np.uint32)
    indices = np.array([1, 2, 3, 0], np.uint32)
    ts, td = oracle.random_walk(indptr, indices, np.array([0, 2], np.uint32), 3, 0.0, 2, 1, 1, 0)
    # layout [node][step][walk] (cuda_sampling_random_walk.cu:60-78)
    assert ts.reshape(2, 3, 2)[0].tolist() == [[0, 0]] * 3 and td.reshape(2, 3, 2)[0].tolist() == [[1, 1], [2, 2], [3, 3]]
    assert td.reshape(2, 3, 2)[1].tolist() == [[3, 3], [0, 0], [1, 1]]
    # restart probability 1: every walk dies after its first step
    ts, td = oracle.random_walk(indptr, indices, np.array([0], np.uint32), 3, 1.0, 2, 1, 1, 0)
    assert ts.reshape(3, 2).tolist() == [[0, 0], [E, E], [E, E]]
    # isolated start node
    ip2 = np.array([0, 0, 1], np.uint32)
    ts, td = oracle.random_walk(ip2, np.array([0], np.uint32), np.array([0], np.uint32), 2, 0.0, 1, 1, 1, 0)
    assert ts.tolist() == [E, E]


# ---- cache / PreSC --------------------------------------------------------------------------------------
def test_cache_table_split_and_presc_rank(oracle):
    freq = np.array([3, 0, 7, 7, 1], np.uint32)
    rank = oracle.presc_rank(freq)
    assert rank.tolist() == [3, 2, 0, 4, 1]           # u64 {freq, id} descending: ties -> larger id first
    assert oracle.num_cached(5, 0.5) == 2 and oracle.num_cached(111059956, 0.25) == 27764989
    table = oracle.cache_table_build(rank, 5, 2)
    assert table.tolist() == [E, E, 1, 0, E]
    ms, md, cs, cd = oracle.cache_split(table, np.array([2, 4, 3, 2, 0], np.uint32))
    assert ms.tolist() == [4, 0] and md.tolist() == [1, 4]
    assert cs.tolist() == [1, 0, 1] and cd.tolist() == [0, 2, 3]
    f2 = freq.copy()
    oracle.freq_count(f2, np.array([1, 1, 4], np.uint32))
    assert f2.tolist() == [3, 2, 7, 7, 2]


def test_shuffle_is_permutation_and_epoch_dependent(oracle):
    train = np.arange(1000, 2000, dtype=np.uint32)
    a, b = oracle.shuffle(train, 5, 0), oracle.shuffle(train, 5, 1)
    assert sorted(a.tolist()) == train.tolist() and not np.array_equal(a, b)
    assert np.array_equal(a, oracle.shuffle(train, 5, 0))
    keys = np.array([oracle.rand_u32(5, 0, 0xFFFF0001, i, 0) for i in range(1000)], np.uint64)
    assert np.array_equal(a, train[np.argsort(keys, kind="stable")])


# ---- live reference build (only where oracle/_ref exists) ------------------------------------------------
def test_live_reference_agrees(oracle, ref):
    ref.set_threads(1)
    rng = np.random.default_rng(9)
    from conftest import small_graph
    indptr, indices = small_graph()
    V = len(indptr) - 1
    seeds = rng.permutation(V)[:200].astype(np.uint32)
    ids = rng.integers(0, V, size=9000).astype(np.uint32)
    oh = oracle.hashtable(20000)
    oh.fill_unique(seeds)
    oh.fill_duplicates(ids)
    for kind in (0, 2):
        rh = ref.hashtable(kind, V if kind == 2 else 20000)
        rh.populate(seeds)
        rh.populate(ids)
        assert np.array_equal(rh.map_nodes(), oh.unique())
        assert np.array_equal(rh.map_edges(ids, ids)[0], oh.map(ids))
    s0, _ = ref.sample_khop0(indptr, indices, seeds, 7)
    so, _ = oracle.sample_khop0(indptr, indices, seeds, 7, 1, 1, 1)
    assert np.array_equal(s0, so)
    feat = rng.standard_normal((V, 9)).astype(np.float32)
    assert np.array_equal(ref.extract(feat, ids), oracle.extract(feat, ids))
    assert ref.predict_num_nodes(8000, [15, 10, 5]) == oracle.predict_num_nodes(8000, [15, 10, 5])


def test_coo_to_csc_oracle_is_scipy_csc(oracle):
    """Block hand-off: the oracle's counting sort by dst == scipy's COO->CSC of the same block
    (DGL's conversion of the block built at adapter.py:92-95), with edge ids as the data to expose the order."""
    sp = pytest.importorskip("scipy.sparse")
    rng = np.random.default_rng(5)
    for e, num_src, num_dst in [(0, 3, 4), (1, 1, 1), (500, 40, 30), (3000, 1000, 512)]:
        row = rng.integers(0, num_src, size=e).astype(np.uint32)
        col = rng.integers(0, num_dst, size=e).astype(np.uint32)
        indptr, indices, eids = oracle.coo_to_csc(row, col, num_dst)
        assert indptr[0] == 0 and indptr[-1] == e and len(indptr) == num_dst + 1
        assert np.array_equal(col[eids], np.repeat(np.arange(num_dst), np.diff(indptr.astype(np.int64))))
        assert np.array_equal(row[eids], indices)
        for d in range(num_dst):                       # stable: edge ids ascend inside every column
            seg = eids[indptr[d]:indptr[d + 1]].astype(np.int64)
            assert (np.diff(seg) > 0).all()
        # scipy sums duplicates, so compare the structure through per-(src,dst) multiplicities
        a = sp.coo_matrix((np.ones(e), (row, col)), shape=(num_src, num_dst)).tocsc()
        b = sp.csc_matrix((np.ones(e), indices.astype(np.int64), indptr.astype(np.int64)), shape=(num_src, num_dst))
        b.sum_duplicates()
        assert (a != b).nnz == 0


# ---------------------------------------------------------------------------------------------
# dataset tools: the oracle's restatements vs the REFERENCE's own offline tools (fixture generated by
# tests/golden/make_golden_tools.py from utility/data-process/toolkit/*, compiled in place)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tools_golden():
    import os
    from fgnn_b200.synth import make_dataset_numpy
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tools_golden.npz"))
    ds = make_dataset_numpy(tuple(int(x) for x in g["spec"]), seed=int(g["seed"]))
    # the fixture is only meaningful for the dataset it was generated from
    assert int(ds["indptr"].astype(np.int64).sum()) == int(g["indptr_sum"])
    assert int(ds["indices"].astype(np.int64).sum()) == int(g["indices_sum"])
    assert int(ds["train_set"].astype(np.int64).sum()) == int(g["train_sum"])
    return g, ds


def policy_weights(ds, policy):
    """create_alias_table.cc:75-92 / create_prob_prefix_table.cc:73-90 (src = the neighbour, indices[off+i])."""
    V = len(ds["indptr"]) - 1
    outdeg = np.bincount(ds["indices"], minlength=V)
    if policy == "kInverseSrcDegreeRand":
        return (1.0 / outdeg[ds["indices"]].astype(np.float64)).astype(np.float32)     # `return 1.0 / src_out_deg;` -> float
    return np.where(outdeg[ds["indices"]] < 10, 100.0, 1.0).astype(np.float32)         # kSrcSuffix


@pytest.mark.parametrize("policy", ["kInverseSrcDegreeRand", "kSrcSuffix"])
def test_weight_tables_match_reference_tools(oracle, tools_golden, policy):
    g, ds = tools_golden
    w = policy_weights(ds, policy)
    prob, alias = oracle.build_alias_table(ds["indptr"], ds["indices"], w)
    assert np.array_equal(prob.view(np.uint32), g["prob_" + policy].view(np.uint32))      # bit-exact fp32
    assert np.array_equal(alias, g["alias_" + policy])
    prefix = oracle.build_prefix_table(ds["indptr"], w)
    assert np.array_equal(prefix.view(np.uint32), g["prefix_" + policy].view(np.uint32))
    # the quirk the restatement keeps: entries whose probability ends at exactly 1 keep a zero alias
    assert (alias[prob == 1.0] == 0).all() and (prob == 1.0).any()


def test_cache_rankings_match_reference_tools(oracle, tools_golden):
    g, ds = tools_golden
    assert np.array_equal(oracle.rank_by_degree(ds["indptr"], ds["indices"]), g["cache_by_degree"])
    assert np.array_equal(oracle.rank_by_heuristic(ds["indptr"], ds["indices"], ds["train_set"]),
                          g["cache_by_heuristic"])
    assert sorted(g["cache_by_heuristic"].tolist()) == list(range(len(ds["indptr"]) - 1))
