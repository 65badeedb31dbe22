"""Property tests (hypothesis) of the CPU oracle against independent plain-Python / numpy models, on random small
inputs including the ragged and empty cases: the oracle is the checker of every GPU parity test, so its own semantics
are pinned three ways — reference fixtures (test_oracle_cpu.py), the reference's code run live, and these models."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

ids = st.lists(st.integers(0, 60), max_size=80)
SET = dict(max_examples=60, deadline=None)


@settings(**SET)
@given(seeds=st.lists(st.integers(0, 60), max_size=20, unique=True), rounds=st.lists(ids, max_size=4))
def test_ordered_hashtable_is_first_occurrence_numbering(oracle, seeds, rounds):
    """OrderedHashTable (cuda_hashtable.cu:131-174,387-438 == CPUHashTable0, cpu_hashtable0.cc:37-47): seeds keep
    local ids 0..S-1, every later id gets the next free local id at its FIRST occurrence."""
    ht = oracle.hashtable(256)
    ht.fill_unique(np.array(seeds, np.uint32))
    model = {s: i for i, s in enumerate(seeds)}
    for r in rounds:
        ht.fill_duplicates(np.array(r, np.uint32))
        for x in r:
            model.setdefault(x, len(model))
        order = sorted(model, key=model.get)
        assert ht.num_items == len(model)
        assert ht.unique().tolist() == order
        if r:
            assert ht.map(np.array(r, np.uint32)).tolist() == [model[x] for x in r]
    ht.reset()
    assert ht.num_items == 0


@settings(**SET)
@given(num_nodes=st.integers(1, 50), pct=st.floats(0, 1), data=st.data())
def test_cache_table_and_split_model(oracle, num_nodes, pct, data):
    """cuda_cache.cu:33-158: table[rank[i]] = i for i < num_cached, stable two-way split of the batch's nodes."""
    rank = np.array(data.draw(st.permutations(range(num_nodes))), np.uint32)
    nodes = np.array(data.draw(st.lists(st.integers(0, num_nodes - 1), max_size=60)), np.uint32)
    nc = oracle.num_cached(num_nodes, pct)
    assert nc == int(num_nodes * pct)                           # cuda_cache_manager_host.cc:66
    table = oracle.cache_table_build(rank, num_nodes, nc)
    exp = np.full(num_nodes, 0xFFFFFFFF, np.uint32)
    exp[rank[:nc]] = np.arange(nc, dtype=np.uint32)
    assert np.array_equal(table, exp)
    ms, md, cs, cd = oracle.cache_split(table, nodes)
    miss = exp[nodes] == 0xFFFFFFFF
    assert np.array_equal(ms, nodes[miss]) and np.array_equal(md, np.nonzero(miss)[0])
    assert np.array_equal(cs, exp[nodes][~miss]) and np.array_equal(cd, np.nonzero(~miss)[0])


@settings(**SET)
@given(freq=st.lists(st.integers(0, 5), min_size=1, max_size=60))
def test_presc_rank_is_descending_freq_then_descending_id(oracle, freq):
    """pre_sampler.cc:44-49,97-99: sort u64 {freq:hi32, id:lo32} descending."""
    f = np.array(freq, np.uint32)
    exp = sorted(range(len(freq)), key=lambda i: (freq[i], i), reverse=True)
    assert oracle.presc_rank(f).tolist() == exp


@settings(**SET)
@given(col=st.lists(st.integers(0, 12), max_size=70), data=st.data())
def test_coo_to_csc_model(oracle, col, data):
    col = np.array(col, np.uint32)
    row = np.array(data.draw(st.lists(st.integers(0, 30), min_size=len(col), max_size=len(col))), np.uint32)
    indptr, indices, eids = oracle.coo_to_csc(row, col, 13)
    perm = np.argsort(col, kind="stable")
    assert np.array_equal(eids, perm.astype(np.uint32)) and np.array_equal(indices, row[perm])
    assert np.array_equal(indptr, np.searchsorted(col[perm], np.arange(14)).astype(np.uint32))


def random_csr(draw, max_nodes=12, max_deg=9):
    V = draw(st.integers(1, max_nodes))
    deg = draw(st.lists(st.integers(0, max_deg), min_size=V, max_size=V))
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    indices = np.array(draw(st.lists(st.integers(0, V - 1), min_size=int(indptr[-1]), max_size=int(indptr[-1]))),
                       np.uint32)
    return V, indptr, indices


@settings(**SET)
@given(data=st.data(), fanout=st.integers(1, 6), which=st.sampled_from(["khop0", "khop2"]))
def test_uniform_samplers_take_min_deg_fanout_distinct_positions(oracle, data, fanout, which):
    """cuda_sampling_khop0.cu / khop2.cu: per seed min(deg, f) edges, seed-major, drawn WITHOUT replacement from the
    row's positions (a multiset inclusion when the row itself repeats a neighbour)."""
    V, indptr, indices = random_csr(data.draw)
    seeds = np.array(data.draw(st.lists(st.integers(0, V - 1), max_size=10)), np.uint32)
    fn = oracle.sample_khop0 if which == "khop0" else oracle.sample_khop2
    src, dst = fn(indptr, indices, seeds, fanout, 7, 3, 1)
    deg = np.diff(indptr.astype(np.int64))
    counts = np.minimum(deg[seeds], fanout) if len(seeds) else np.zeros(0, np.int64)
    assert np.array_equal(src, np.repeat(seeds, counts))
    o = 0
    for s, c in zip(seeds, counts):
        picked = sorted(dst[o:o + c].tolist())
        row = sorted(indices[indptr[s]:indptr[s + 1]].tolist())
        if c == deg[s]:
            assert picked == row                                 # short rows are copied whole
        else:
            it = iter(row)                                        # multiset inclusion
            assert all(any(x == y for y in it) for x in picked)
        o += c
    # same counters -> same draws
    src2, dst2 = fn(indptr, indices, seeds, fanout, 7, 3, 1)
    assert np.array_equal(dst, dst2)


@settings(**SET)
@given(data=st.data(), fanout=st.integers(1, 6))
def test_khop1_is_sorted_by_seed_id_with_adjacent_dedup(oracle, data, fanout):
    """cuda_sampling_khop1.cu:84-86,169-178: with replacement, stable sort by seed id, equal neighbours merged only
    when adjacent."""
    V, indptr, indices = random_csr(data.draw)
    seeds = np.array(data.draw(st.lists(st.integers(0, V - 1), max_size=8, unique=True)), np.uint32)
    src, dst = oracle.sample_khop1(indptr, indices, seeds, fanout, 11, 5, 0)
    assert np.array_equal(src, np.sort(src, kind="stable"))
    deg = np.diff(indptr.astype(np.int64))
    for s in set(src.tolist()):
        assert deg[s] > 0
        mine = dst[src == s]
        assert 1 <= len(mine) <= fanout
        assert all(a != b for a, b in zip(mine[:-1], mine[1:]))      # no adjacent duplicates survive
        assert set(mine.tolist()) <= set(indices[indptr[s]:indptr[s + 1]].tolist())
    assert set(src.tolist()) == {int(s) for s in seeds if deg[s] > 0}


@settings(**SET)
@given(data=st.data(), K=st.integers(1, 4), per_node=st.integers(1, 8))
def test_topk_model(oracle, data, K, per_node):
    """cuda_frequency_hashmap.cu:361-401,502-504,585-676: per start node, multiplicities of the visited ids, top K by
    count, ties in first-occurrence order, dead records (EMPTY start) ignored."""
    n = data.draw(st.integers(0, 5))
    inp = np.arange(100, 100 + n, dtype=np.uint32)
    E = 0xFFFFFFFF
    ts, td, exp = [], [], []
    for node in inp:
        visits = data.draw(st.lists(st.one_of(st.none(), st.integers(0, 5)), min_size=per_node, max_size=per_node))
        first, count = {}, {}
        for p, v in enumerate(visits):
            ts.append(E if v is None else int(node))
            td.append(0 if v is None else v)
            if v is not None:
                first.setdefault(v, p)
                count[v] = count.get(v, 0) + 1
        for v in sorted(count, key=lambda v: (-count[v], first[v]))[:K]:
            exp.append((int(node), v, count[v]))
    s, d, c = oracle.topk(np.array(ts, np.uint32), np.array(td, np.uint32), inp, per_node, K)
    assert list(zip(s.tolist(), d.tolist(), c.tolist())) == exp


@settings(**SET)
@given(data=st.data(), walk_len=st.integers(1, 4), num_walk=st.integers(1, 3), p=st.sampled_from([0.0, 0.3, 1.0]))
def test_random_walk_follows_edges_and_stays_dead(oracle, data, walk_len, num_walk, p):
    """cuda_sampling_random_walk.cu:43-109: layout [node][step][walk]; step 0 leaves from the start node, every later
    step from the previous step's node; a walk that died (restart or dead end) stays dead; records carry the START
    node as src."""
    V, indptr, indices = random_csr(data.draw)
    starts = np.array(data.draw(st.lists(st.integers(0, V - 1), max_size=5)), np.uint32)
    ts, td = oracle.random_walk(indptr, indices, starts, walk_len, p, num_walk, 3, 9, 2)
    E = 0xFFFFFFFF
    ts = ts.reshape(len(starts), walk_len, num_walk)
    td = td.reshape(len(starts), walk_len, num_walk)
    for n, s0 in enumerate(starts):
        for w in range(num_walk):
            cur, alive = int(s0), True
            for s in range(walk_len):
                if ts[n, s, w] == E:
                    alive = False
                    continue
                assert alive, "a dead walk came back to life"
                assert ts[n, s, w] == s0
                assert td[n, s, w] in indices[indptr[cur]:indptr[cur + 1]]
                cur = int(td[n, s, w])
            if p == 1.0 and walk_len > 1:
                assert (ts[n, 1:, w] == E).all()
