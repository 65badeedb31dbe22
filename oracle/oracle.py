"""TEST INFRASTRUCTURE — ctypes front-end to the CPU oracle.

`Oracle`   : our C restatement (oracle/fgnn_oracle.c -> _build/libfgnn_oracle.so)
`RefCPU`   : the reference's own CPU translation units (oracle/_ref/, built from
             /root/reference by `make -C oracle ref`; may be absent)
`sample_batch_oracle` : numpy restatement of the per-batch driver loop
             DoGPUSample (samgraph/common/cuda/cuda_loops.cc:50-267).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product path never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMPTY = 0xFFFFFFFF

u32p = C.POINTER(C.c_uint32)
f32p = C.POINTER(C.c_float)
szp = C.POINTER(C.c_size_t)


def _p(a, ty=u32p):
    if a is None:
        return None
    return a.ctypes.data_as(ty)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def build_oracle(force=False):
    so = os.path.join(HERE, "_build", "libfgnn_oracle.so")
    src = os.path.join(HERE, "fgnn_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return so


def build_ref():
    """Build oracle/_ref from /root/reference when that tree exists."""
    so = os.path.join(HERE, "_ref", "libsamgraph_ref_cpu.so")
    if os.path.isdir(os.environ.get("FGNN_REFERENCE", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return so if os.path.exists(so) else None


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build_oracle())
        L = self.lib
        L.fgo_rand_u32.restype = C.c_uint32
        L.fgo_rand_u32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.fgo_uniform_f32.restype = C.c_float
        L.fgo_uniform_f32.argtypes = [C.c_uint32]
        L.fgo_uniform_f64.restype = C.c_double
        L.fgo_uniform_f64.argtypes = [C.c_uint32, C.c_uint32]
        L.fgo_predict_num_nodes.restype = C.c_size_t
        L.fgo_predict_num_nodes.argtypes = [C.c_size_t, szp, C.c_size_t]
        L.fgo_table_size.restype = C.c_size_t
        L.fgo_table_size.argtypes = [C.c_size_t, C.c_size_t]
        common = [u32p, u32p, u32p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p]
        for name in ("fgo_sample_khop0", "fgo_sample_khop2", "fgo_sample_khop1"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = common
        L.fgo_sample_weighted_khop.restype = C.c_size_t
        L.fgo_sample_weighted_khop.argtypes = [u32p, u32p, f32p, u32p, u32p, C.c_size_t, C.c_size_t,
                                               C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p]
        L.fgo_sample_weighted_khop_hash_dedup.restype = C.c_size_t
        L.fgo_sample_weighted_khop_hash_dedup.argtypes = L.fgo_sample_weighted_khop.argtypes
        L.fgo_sample_weighted_khop_prefix.restype = C.c_size_t
        L.fgo_sample_weighted_khop_prefix.argtypes = [u32p, u32p, f32p, u32p, C.c_size_t, C.c_size_t,
                                                      C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p]
        L.fgo_random_walk.restype = None
        L.fgo_random_walk.argtypes = [u32p, u32p, u32p, C.c_size_t, C.c_size_t, C.c_double, C.c_size_t,
                                      C.c_uint64, C.c_uint64, C.c_uint32, u32p, u32p]
        L.fgo_topk.restype = C.c_size_t
        L.fgo_topk.argtypes = [u32p, u32p, u32p, C.c_size_t, C.c_size_t, C.c_size_t, u32p, u32p, u32p]
        L.fgo_hashtable_new.restype = C.c_void_p
        L.fgo_hashtable_new.argtypes = [C.c_size_t]
        L.fgo_hashtable_free.argtypes = [C.c_void_p]
        L.fgo_hashtable_reset.argtypes = [C.c_void_p]
        L.fgo_hashtable_num_items.restype = C.c_size_t
        L.fgo_hashtable_num_items.argtypes = [C.c_void_p]
        L.fgo_hashtable_fill_unique.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.fgo_hashtable_fill_duplicates.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.fgo_hashtable_unique.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.fgo_hashtable_map.restype = C.c_int
        L.fgo_hashtable_map.argtypes = [C.c_void_p, u32p, C.c_size_t, u32p]
        L.fgo_num_cached.restype = C.c_size_t
        L.fgo_num_cached.argtypes = [C.c_size_t, C.c_double]
        L.fgo_cache_table_build.argtypes = [u32p, C.c_size_t, C.c_size_t, u32p]
        L.fgo_cache_split.argtypes = [u32p, u32p, C.c_size_t, u32p, u32p, szp, u32p, u32p, szp]
        L.fgo_row_copy.argtypes = [C.c_void_p, u32p, C.c_void_p, u32p, C.c_size_t, C.c_size_t, C.c_uint64]
        L.fgo_freq_count.argtypes = [u32p, u32p, C.c_size_t]
        L.fgo_presc_rank.argtypes = [u32p, C.c_size_t, u32p]
        L.fgo_build_alias_table.argtypes = [u32p, u32p, C.c_size_t, f32p, f32p, u32p]
        L.fgo_build_prefix_table.argtypes = [u32p, C.c_size_t, f32p, f32p]
        L.fgo_set_threads.argtypes = [C.c_int]
        L.fgo_shuffle.argtypes = [u32p, C.c_size_t, C.c_uint64, C.c_uint64, u32p]

    # ---- rng ----
    def rand_u32(self, seed, batch_key, tag, item, draw):
        return int(self.lib.fgo_rand_u32(seed, batch_key, tag, item, draw))

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.fgo_philox4x32_10(c, k, o)
        return [int(x) for x in o]

    def predict_num_nodes(self, batch, fanout):
        f = (C.c_size_t * len(fanout))(*fanout)
        return int(self.lib.fgo_predict_num_nodes(batch, f, len(fanout)))

    def shuffle(self, train_set, seed, epoch):
        train_set = _u32(train_set)
        out = np.empty(len(train_set), np.uint32)
        self.lib.fgo_shuffle(_p(train_set), len(train_set), seed, epoch, _p(out))
        return out

    def table_size(self, num, scale=2):
        return int(self.lib.fgo_table_size(num, scale))

    # ---- samplers ----
    def _khop(self, fn, indptr, indices, inp, fanout, seed, batch_key, tag):
        inp = _u32(inp)
        n = len(inp)
        src = np.empty(n * fanout + 1, np.uint32)
        dst = np.empty(n * fanout + 1, np.uint32)
        m = fn(_p(indptr), _p(indices), _p(inp), n, fanout, seed, batch_key, tag, _p(src), _p(dst))
        return src[:m].copy(), dst[:m].copy()

    def sample_khop0(self, indptr, indices, inp, fanout, seed, batch_key, tag):
        return self._khop(self.lib.fgo_sample_khop0, indptr, indices, inp, fanout, seed, batch_key, tag)

    def sample_khop2(self, indptr, indices, inp, fanout, seed, batch_key, tag):
        return self._khop(self.lib.fgo_sample_khop2, indptr, indices, inp, fanout, seed, batch_key, tag)

    def sample_khop1(self, indptr, indices, inp, fanout, seed, batch_key, tag):
        return self._khop(self.lib.fgo_sample_khop1, indptr, indices, inp, fanout, seed, batch_key, tag)

    def sample_weighted_khop(self, indptr, indices, prob, alias, inp, fanout, seed, batch_key, tag,
                             hash_dedup=False):
        inp = _u32(inp)
        n = len(inp)
        src = np.empty(n * fanout + 1, np.uint32)
        dst = np.empty(n * fanout + 1, np.uint32)
        fn = self.lib.fgo_sample_weighted_khop_hash_dedup if hash_dedup else self.lib.fgo_sample_weighted_khop
        m = fn(_p(indptr), _p(indices), _p(prob, f32p), _p(alias), _p(inp), n, fanout, seed, batch_key,
               tag, _p(src), _p(dst))
        return src[:m].copy(), dst[:m].copy()

    def sample_weighted_khop_prefix(self, indptr, indices, prefix, inp, fanout, seed, batch_key, tag):
        inp = _u32(inp)
        n = len(inp)
        src = np.empty(n * fanout + 1, np.uint32)
        dst = np.empty(n * fanout + 1, np.uint32)
        m = self.lib.fgo_sample_weighted_khop_prefix(_p(indptr), _p(indices), _p(prefix, f32p), _p(inp), n,
                                                     fanout, seed, batch_key, tag, _p(src), _p(dst))
        return src[:m].copy(), dst[:m].copy()

    def random_walk(self, indptr, indices, inp, walk_len, restart_prob, num_walk, seed, batch_key, tag):
        inp = _u32(inp)
        n = len(inp)
        ts = np.empty(n * walk_len * num_walk + 1, np.uint32)
        td = np.empty(n * walk_len * num_walk + 1, np.uint32)
        self.lib.fgo_random_walk(_p(indptr), _p(indices), _p(inp), n, walk_len, restart_prob, num_walk,
                                 seed, batch_key, tag, _p(ts), _p(td))
        return ts[:-1], td[:-1]

    def topk(self, tmp_src, tmp_dst, inp, edges_per_node, K):
        inp = _u32(inp)
        n = len(inp)
        s = np.empty(n * K + 1, np.uint32)
        d = np.empty(n * K + 1, np.uint32)
        w = np.empty(n * K + 1, np.uint32)
        m = self.lib.fgo_topk(_p(_u32(tmp_src)), _p(_u32(tmp_dst)), _p(inp), n, edges_per_node, K,
                              _p(s), _p(d), _p(w))
        return s[:m].copy(), d[:m].copy(), w[:m].copy()

    def sample_random_walk(self, indptr, indices, inp, walk_len, restart_prob, num_walk, K, seed,
                           batch_key, tag):
        ts, td = self.random_walk(indptr, indices, inp, walk_len, restart_prob, num_walk, seed, batch_key, tag)
        return self.topk(ts, td, inp, walk_len * num_walk, K)

    # ---- hashtable ----
    def hashtable(self, max_items):
        return OracleHashTable(self, max_items)

    # ---- cache ----
    def num_cached(self, num_nodes, pct):
        return int(self.lib.fgo_num_cached(num_nodes, pct))

    def cache_table_build(self, rank, num_nodes, num_cached):
        table = np.empty(num_nodes, np.uint32)
        self.lib.fgo_cache_table_build(_p(_u32(rank)), num_nodes, num_cached, _p(table))
        return table

    def cache_split(self, table, nodes):
        nodes = _u32(nodes)
        n = len(nodes)
        ms, md, cs, cd = (np.empty(n + 1, np.uint32) for _ in range(4))
        nm, nc = C.c_size_t(0), C.c_size_t(0)
        self.lib.fgo_cache_split(_p(table), _p(nodes), n, _p(ms), _p(md), C.byref(nm), _p(cs), _p(cd),
                                 C.byref(nc))
        return ms[:nm.value].copy(), md[:nm.value].copy(), cs[:nc.value].copy(), cd[:nc.value].copy()

    def row_copy(self, dst, dst_index, src, src_index, n, row_bytes, mask=0xFFFFFFFFFFFFFFFF):
        self.lib.fgo_row_copy(dst.ctypes.data, _p(dst_index), src.ctypes.data, _p(src_index), n, row_bytes,
                              mask)

    def extract(self, src, index, mask=0xFFFFFFFFFFFFFFFF):
        index = _u32(index)
        src = np.ascontiguousarray(src)
        row = src[0:1].nbytes if src.ndim > 1 else src.itemsize
        out = np.empty((len(index),) + src.shape[1:], src.dtype)
        self.row_copy(out, None, src, index, len(index), row, mask)
        return out

    # ---- presc ----
    def freq_count(self, freq, nodes):
        nodes = _u32(nodes)
        self.lib.fgo_freq_count(_p(freq), _p(nodes), len(nodes))

    def presc_rank(self, freq):
        rank = np.empty(len(freq), np.uint32)
        self.lib.fgo_presc_rank(_p(_u32(freq)), len(freq), _p(rank))
        return rank

    # ---- weight tables ----
    def build_alias_table(self, indptr, indices, weights):
        prob = np.zeros(len(indices), np.float32)
        alias = np.zeros(len(indices), np.uint32)
        self.lib.fgo_build_alias_table(_p(indptr), _p(indices), len(indptr) - 1,
                                       _p(np.ascontiguousarray(weights, np.float32), f32p),
                                       _p(prob, f32p), _p(alias))
        return prob, alias

    def build_prefix_table(self, indptr, weights):
        pre = np.zeros(len(weights), np.float32)
        self.lib.fgo_build_prefix_table(_p(indptr), len(indptr) - 1,
                                        _p(np.ascontiguousarray(weights, np.float32), f32p), _p(pre, f32p))
        return pre

    # ---- cache rankings of the offline tools ----
    def rank_by_degree(self, indptr, indices):
        """toolkit/cache/cache_by_degree.cc:29-62: sort {out_degree, id} pairs descending (ties: larger id first);
        out_degree[v] = occurrences of v in `indices` (common/graph_loader.cc:109-147)."""
        V = len(indptr) - 1
        return self.presc_rank(np.bincount(_u32(indices), minlength=V).astype(np.uint32))

    def rank_by_heuristic(self, indptr, indices, train_set):
        """toolkit/cache/cache_by_heuristic.cc:28-91: (1) the training nodes in train_set order, (2) their first-hop
        neighbours in order of first appearance (train_set order, CSR order inside a row), (3) every other vertex
        by {out_degree, id} descending."""
        V = len(indptr) - 1
        added = np.zeros(V, bool)
        rank = []
        for t in _u32(train_set):
            rank.append(int(t))
            added[t] = True                       # (the tool assumes train_set has no duplicates)
        for t in _u32(train_set):
            for nb in indices[indptr[t]:indptr[t + 1]]:
                if not added[nb]:
                    rank.append(int(nb))
                    added[nb] = True
        for v in self.rank_by_degree(indptr, indices):
            if not added[v]:
                rank.append(int(v))
                added[v] = True
        return np.array(rank, np.uint32)

    def rank_by_degree_hop(self, indptr, indices, train_set, hops=2):
        """toolkit/cache/cache_by_degree_hop.cc:31-180: vertices within `hops` hops of the training set (hopNodes,
        :31-82) are ranked first, by their out-degree inside the sub-graph that keeps only those vertices' adjacency
        rows (gen_khop_graph :85-118, merged with bit 30 set :120-131); all others by whole-graph out-degree."""
        indptr, indices = _u32(indptr), _u32(indices)
        V = len(indptr) - 1
        before = np.zeros(V, np.uint8)
        before[_u32(train_set)] = 2
        for _ in range(hops):
            after = np.zeros(V, bool)
            for v in np.nonzero(before == 2)[0]:
                after[indices[indptr[v]:indptr[v + 1]]] = True
            fresh = after & (before == 0)
            before[before != 0] = 1
            before[fresh] = 2
        touched = before != 0
        whole = np.bincount(indices, minlength=V).astype(np.uint32)
        keep = np.repeat(touched, np.diff(indptr.astype(np.int64)))
        sub = np.bincount(indices[keep], minlength=V).astype(np.uint32)
        return self.presc_rank(np.where(touched, sub | np.uint32(0x40000000), whole).astype(np.uint32))

    @staticmethod
    def rank_by_fake_optimal(indptr, indices, train_set, fanout=(25, 10), order_threads=1):
        """toolkit/cache/cache_by_fake_optimal.cc:61-183 with the tool's batch_size = 1 (:173): per training node,
        hop1_miss[v] = prod over its edges to v of max(0, 1 - fanout[1]/deg(seed)); hop2_miss[w] = prod over the
        edges h->w of the touched vertices h of (1 - (1 - hop1_miss[h]) * min(1, fanout[0]/deg(h))), the touched
        vertices visited in TouchedNodeCtx::compact() order (buckets id % threads, :44-60); the seed itself is a
        certain hit; expectation[v] += 1 - hop1_miss[v]*hop2_miss[v]; ranking = {expectation, id} descending.
        Plain IEEE doubles in the tool's operation order, so the ranking is bit-exact against the tool."""
        indptr, indices = _u32(indptr), _u32(indices)
        V, T = len(indptr) - 1, int(order_threads)
        exp = [0.0] * V
        for t in (int(x) for x in _u32(train_set)):
            buckets = [[] for _ in range(T)]
            seen = set()

            def touch(v):
                if v not in seen:
                    seen.add(v)
                    buckets[v % T].append(v)
            touch(t)
            h1, h2 = {}, {}
            row = [int(x) for x in indices[indptr[t]:indptr[t + 1]]]
            if row:
                miss = max(0.0, 1 - fanout[1] / float(len(row)))
                for d in row:
                    h1[d] = h1.get(d, 1.0) * miss
                    touch(d)
            h1[t] = 0.0
            for h in [v for b in buckets for v in b]:
                nb = [int(x) for x in indices[indptr[h]:indptr[h + 1]]]
                if not nb:
                    continue
                b1_hit = 1 - h1.get(h, 1.0)
                b2_hit = min(1.0, fanout[0] / float(len(nb)))
                path_miss = 1 - b1_hit * b2_hit
                for w in nb:
                    h2[w] = h2.get(w, 1.0) * path_miss
                    touch(w)
            h2[t] = 0.0
            for v in seen:
                a, b = h1.get(v, 1.0), h2.get(v, 1.0)
                if a == 1 and b == 1:
                    continue
                exp[v] += 1 - a * b
        e = np.array(exp, np.float64)
        order = np.lexsort((np.arange(V), e))[::-1]          # {expectation, id} descending (std::greater on pairs)
        return order.astype(np.uint32), e

    # ---- block hand-off ----
    @staticmethod
    def coo_to_csc(row, col, num_dst):
        """CSC arrays of one sampled layer as the reference's DGL patch consumes them
        (3rdparty/dgl.patch:30-57 create_unitgraph_from_csc(indptr, indices, edge_ids)); the block itself is
        built from (row, col) = (src, dst) at samgraph/torch/adapter.py:92-95.  Plain counting sort by dst,
        stable in the edge id — what DGL's COO->CSC conversion produces for these blocks."""
        row, col = _u32(row), _u32(col)
        counts = np.zeros(num_dst + 1, np.int64)
        for c in col:                      # histogram of dst
            counts[int(c) + 1] += 1
        indptr = np.cumsum(counts)
        fill = indptr[:-1].copy()
        indices = np.empty(len(row), np.uint32)
        eids = np.empty(len(row), np.uint32)
        for e in range(len(row)):          # stable placement
            k = fill[int(col[e])]
            indices[k], eids[k] = row[e], e
            fill[int(col[e])] += 1
        return indptr.astype(np.uint32), indices, eids


class OracleHashTable:
    def __init__(self, oracle, max_items):
        self.o = oracle
        self.h = C.c_void_p(oracle.lib.fgo_hashtable_new(max_items))

    def __del__(self):
        try:
            self.o.lib.fgo_hashtable_free(self.h)
        except Exception:
            pass

    def reset(self):
        self.o.lib.fgo_hashtable_reset(self.h)

    @property
    def num_items(self):
        return int(self.o.lib.fgo_hashtable_num_items(self.h))

    def fill_unique(self, ids):
        ids = _u32(ids)
        self.o.lib.fgo_hashtable_fill_unique(self.h, _p(ids), len(ids))

    def fill_duplicates(self, ids):
        ids = _u32(ids)
        self.o.lib.fgo_hashtable_fill_duplicates(self.h, _p(ids), len(ids))
        return self.unique()

    def unique(self):
        out = np.empty(self.num_items, np.uint32)
        self.o.lib.fgo_hashtable_unique(self.h, _p(out), len(out))
        return out

    def map(self, ids):
        ids = _u32(ids)
        out = np.empty(len(ids), np.uint32)
        rc = self.o.lib.fgo_hashtable_map(self.h, _p(ids), len(ids), _p(out))
        assert rc == 0, "id missing from ordered hash table"
        return out


class RefCPU:
    """The reference's own CPU code (oracle/_ref)."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "_ref", "libsamgraph_ref_cpu.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L = self.lib
        for name in ("ref_cpu_sample_khop0", "ref_cpu_sample_khop2"):
            getattr(L, name).argtypes = [u32p, u32p, u32p, C.c_size_t, u32p, u32p, szp, C.c_size_t]
        L.ref_hashtable_new.restype = C.c_void_p
        L.ref_hashtable_new.argtypes = [C.c_int, C.c_size_t]
        L.ref_hashtable_free.argtypes = [C.c_void_p]
        L.ref_hashtable_populate.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.ref_hashtable_map_nodes.argtypes = [C.c_void_p, u32p, C.c_size_t]
        L.ref_hashtable_map_edges.argtypes = [C.c_void_p, u32p, u32p, C.c_size_t, u32p, u32p]
        L.ref_hashtable_reset.argtypes = [C.c_void_p]
        L.ref_hashtable_num_items.restype = C.c_size_t
        L.ref_hashtable_num_items.argtypes = [C.c_void_p]
        L.ref_cpu_extract.argtypes = [C.c_void_p, C.c_void_p, u32p, C.c_size_t, C.c_size_t, C.c_int]
        L.ref_predict_num_nodes.restype = C.c_size_t
        L.ref_predict_num_nodes.argtypes = [C.c_size_t, szp, C.c_size_t]
        L.ref_set_omp_threads.argtypes = [C.c_int]
        L.ref_get_omp_threads.restype = C.c_int

    def set_threads(self, n):
        self.lib.ref_set_omp_threads(n)

    def _khop(self, fn, indptr, indices, inp, fanout):
        inp = _u32(inp)
        n = len(inp)
        src = np.empty(n * fanout + 1, np.uint32)
        dst = np.empty(n * fanout + 1, np.uint32)
        m = C.c_size_t(0)
        fn(_p(indptr), _p(indices), _p(inp), n, _p(src), _p(dst), C.byref(m), fanout)
        return src[:m.value].copy(), dst[:m.value].copy()

    def sample_khop0(self, indptr, indices, inp, fanout):
        return self._khop(self.lib.ref_cpu_sample_khop0, indptr, indices, inp, fanout)

    def sample_khop2(self, indptr, indices_mutable, inp, fanout):
        return self._khop(self.lib.ref_cpu_sample_khop2, indptr, indices_mutable, inp, fanout)

    def hashtable(self, kind, max_items):
        return RefHashTable(self, kind, max_items)

    _DT = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int16): 2, np.dtype(np.uint8): 3,
           np.dtype(np.int32): 4, np.dtype(np.int64): 6}

    def extract(self, src, index):
        index = _u32(index)
        src = np.ascontiguousarray(src)
        dim = int(np.prod(src.shape[1:])) if src.ndim > 1 else 1
        out = np.empty((len(index),) + src.shape[1:], src.dtype)
        self.lib.ref_cpu_extract(out.ctypes.data, src.ctypes.data, _p(index), len(index), dim,
                                 self._DT[src.dtype])
        return out

    def predict_num_nodes(self, batch, fanout):
        f = (C.c_size_t * len(fanout))(*fanout)
        return int(self.lib.ref_predict_num_nodes(batch, f, len(fanout)))


class RefHashTable:
    def __init__(self, ref, kind, max_items):
        self.r = ref
        self.h = C.c_void_p(ref.lib.ref_hashtable_new(kind, max_items))

    def __del__(self):
        try:
            self.r.lib.ref_hashtable_free(self.h)
        except Exception:
            pass

    def reset(self):
        self.r.lib.ref_hashtable_reset(self.h)

    @property
    def num_items(self):
        return int(self.r.lib.ref_hashtable_num_items(self.h))

    def populate(self, ids):
        ids = _u32(ids)
        self.r.lib.ref_hashtable_populate(self.h, _p(ids), len(ids))

    def map_nodes(self):
        out = np.empty(self.num_items, np.uint32)
        self.r.lib.ref_hashtable_map_nodes(self.h, _p(out), len(out))
        return out

    def map_edges(self, src, dst):
        src, dst = _u32(src), _u32(dst)
        ns, nd = np.empty(len(src), np.uint32), np.empty(len(dst), np.uint32)
        self.r.lib.ref_hashtable_map_edges(self.h, _p(src), _p(dst), len(src), _p(ns), _p(nd))
        return ns, nd


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "libsamgraph_ref_cpu.so"))


# ---------------------------------------------------------------------------
# per-batch driver: restatement of DoGPUSample (cuda_loops.cc:50-267)
# ---------------------------------------------------------------------------
def sample_batch_oracle(o, graph, seeds, fanouts, sample_type, seed, batch_key, rw=None):
    """graph: dict(indptr, indices[, prob_table, alias_table, prob_prefix_table]).
    fanouts: list as given to samgraph (`fanout`), sampled from last to first.
    Returns dict(layers=[{row,col,data,num_src,num_dst,num_edge}], input_nodes,
    output_nodes, raw=[(src,dst)])."""
    indptr, indices = graph["indptr"], graph["indices"]
    num_layers = len(fanouts)
    ht = o.hashtable(max(16, o.predict_num_nodes(len(seeds), fanouts)))
    ht.fill_unique(seeds)                                # cuda_loops.cc:67-69
    cur = _u32(seeds)
    layers = [None] * num_layers
    raw = [None] * num_layers
    for i in range(num_layers - 1, -1, -1):              # :87
        f = int(fanouts[i])
        data = None
        if sample_type == "khop0":
            s, d = o.sample_khop0(indptr, indices, cur, f, seed, batch_key, i)
        elif sample_type == "khop2":
            s, d = o.sample_khop2(indptr, indices, cur, f, seed, batch_key, i)
        elif sample_type == "khop1":
            s, d = o.sample_khop1(indptr, indices, cur, f, seed, batch_key, i)
        elif sample_type == "weighted_khop":
            s, d = o.sample_weighted_khop(indptr, indices, graph["prob_table"], graph["alias_table"], cur, f,
                                          seed, batch_key, i)
        elif sample_type == "weighted_khop_hash_dedup":
            s, d = o.sample_weighted_khop(indptr, indices, graph["prob_table"], graph["alias_table"], cur, f,
                                          seed, batch_key, i, hash_dedup=True)
        elif sample_type == "weighted_khop_prefix":
            s, d = o.sample_weighted_khop_prefix(indptr, indices, graph["prob_prefix_table"], cur, f, seed,
                                                 batch_key, i)
        elif sample_type == "random_walk":
            s, d, data = o.sample_random_walk(indptr, indices, cur, rw["random_walk_length"],
                                              rw["random_walk_restart_prob"], rw["num_random_walk"],
                                              rw["num_neighbor"], seed, batch_key, i)
        else:
            raise ValueError(sample_type)
        raw[i] = (s, d)
        unique = ht.fill_duplicates(d)                   # :176-186
        new_src = ht.map(s)                              # :203-205
        new_dst = ht.map(d)
        layers[i] = dict(row=new_dst, col=new_src, data=data,          # :210-221
                         num_src=len(unique), num_dst=len(cur), num_edge=len(s))
        cur = unique
    return dict(layers=layers, input_nodes=cur, output_nodes=_u32(seeds), raw=raw)


# ---------------------------------------------------------------------------
# the reference's own CUDA kernels (oracle/_ref/libsamgraph_ref_cuda.so, built by `make -C oracle refcuda` from
# the reference's .cu files in place): GPU-side checker for tests/test_vs_reference_cuda_gpu.py.  All array
# arguments are torch CUDA tensors (int32-typed views of uint32 data, float32 tables).
# ---------------------------------------------------------------------------
def have_ref_cuda():
    return os.path.exists(os.path.join(HERE, "_ref", "libsamgraph_ref_cuda.so"))


class RefCUDA:
    KINDS = {"khop0": 0, "khop1": 1, "weighted_khop": 2, "random_walk": 3, "weighted_khop_prefix": 4, "khop2": 5,
             "weighted_khop_hash_dedup": 6}

    def __init__(self, device=0, path=None):
        path = path or os.path.join(HERE, "_ref", "libsamgraph_ref_cuda.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        L.refcuda_set_device.argtypes = [C.c_int]
        L.refcuda_states_new.restype = vp
        L.refcuda_states_new.argtypes = [C.c_int, szp, sz, sz, sz]
        L.refcuda_states_free.argtypes = [vp]
        L.refcuda_sample.restype = sz
        L.refcuda_sample.argtypes = [C.c_int, vp, vp, vp, vp, vp, vp, sz, sz, vp, vp, vp]
        L.refcuda_freqmap_new.restype = vp
        L.refcuda_freqmap_new.argtypes = [sz, sz]
        L.refcuda_freqmap_free.argtypes = [vp]
        L.refcuda_random_walk.restype = sz
        L.refcuda_random_walk.argtypes = [vp, vp, vp, sz, sz, C.c_double, sz, sz, vp, vp, vp, vp, vp]
        L.refcuda_topk.restype = sz
        L.refcuda_topk.argtypes = [vp, vp, vp, sz, vp, sz, sz, vp, vp, vp]
        L.refcuda_ht_new.restype = vp
        L.refcuda_ht_new.argtypes = [sz]
        L.refcuda_ht_free.argtypes = [vp]
        L.refcuda_ht_reset.argtypes = [vp]
        L.refcuda_ht_fill_unique.argtypes = [vp, vp, sz]
        L.refcuda_ht_fill_duplicates.restype = sz
        L.refcuda_ht_fill_duplicates.argtypes = [vp, vp, sz, vp]
        L.refcuda_ht_num_items.restype = sz
        L.refcuda_ht_num_items.argtypes = [vp]
        L.refcuda_map_edges.argtypes = [vp, vp, vp, vp, vp, sz]
        L.refcuda_get_miss_cache_index.argtypes = [vp, vp, vp, szp, vp, vp, szp, vp, sz]
        L.refcuda_set_device(device)

    @staticmethod
    def _d(t):
        return None if t is None else t.data_ptr()

    def states(self, sample_type, fanouts, batch_size, num_random_walk=1):
        f = (C.c_size_t * len(fanouts))(*fanouts)
        return self.lib.refcuda_states_new(self.KINDS[sample_type], f, len(fanouts), batch_size, num_random_walk)

    def states_free(self, s):
        self.lib.refcuda_states_free(s)

    def sample(self, sample_type, indptr, indices, inp, fanout, states, prob=None, alias=None, prefix=None):
        """One layer of the reference sampler; returns (src, dst) device tensors trimmed to the edge count."""
        import torch
        n = inp.numel()
        out_src = torch.empty(max(1, n * fanout), dtype=torch.int32, device=inp.device)
        out_dst = torch.empty_like(out_src)
        m = self.lib.refcuda_sample(self.KINDS[sample_type], self._d(indptr), self._d(indices), self._d(prob),
                                    self._d(alias), self._d(prefix), self._d(inp), n, fanout, self._d(out_src),
                                    self._d(out_dst), states)
        return out_src[:m], out_dst[:m]

    def freqmap(self, max_nodes, edges_per_node):
        return self.lib.refcuda_freqmap_new(max_nodes, edges_per_node)

    def freqmap_free(self, m):
        self.lib.refcuda_freqmap_free(m)

    def random_walk(self, indptr, indices, inp, walk_len, restart_prob, num_walk, K, freqmap, states):
        import torch
        n = inp.numel()
        outs = [torch.empty(max(1, n * K), dtype=torch.int32, device=inp.device) for _ in range(3)]
        m = self.lib.refcuda_random_walk(self._d(indptr), self._d(indices), self._d(inp), n, walk_len, restart_prob,
                                         num_walk, K, self._d(outs[0]), self._d(outs[1]), self._d(outs[2]), freqmap,
                                         states)
        return [o[:m] for o in outs]

    def topk(self, freqmap, tmp_src, tmp_dst, inp, K):
        import torch
        n = inp.numel()
        outs = [torch.empty(max(1, n * K), dtype=torch.int32, device=inp.device) for _ in range(3)]
        m = self.lib.refcuda_topk(freqmap, self._d(tmp_src), self._d(tmp_dst), tmp_src.numel(), self._d(inp), n, K,
                                  self._d(outs[0]), self._d(outs[1]), self._d(outs[2]))
        return [o[:m] for o in outs]

    def hashtable(self, size):
        return RefCudaHashTable(self, size)

    def get_miss_cache_index(self, table, nodes):
        import torch
        n = nodes.numel()
        outs = [torch.empty(max(1, n), dtype=torch.int32, device=nodes.device) for _ in range(4)]
        nm, nc = C.c_size_t(0), C.c_size_t(0)
        self.lib.refcuda_get_miss_cache_index(self._d(table), self._d(outs[0]), self._d(outs[1]), C.byref(nm),
                                              self._d(outs[2]), self._d(outs[3]), C.byref(nc), self._d(nodes), n)
        return outs[0][:nm.value], outs[1][:nm.value], outs[2][:nc.value], outs[3][:nc.value]


class RefCudaHashTable:
    def __init__(self, ref, size):
        self.r, self.h = ref, ref.lib.refcuda_ht_new(size)

    def __del__(self):
        if getattr(self, "h", None):
            self.r.lib.refcuda_ht_free(self.h)
            self.h = None

    def reset(self):
        self.r.lib.refcuda_ht_reset(self.h)

    def fill_unique(self, ids):
        self.r.lib.refcuda_ht_fill_unique(self.h, ids.data_ptr(), ids.numel())

    def fill_duplicates(self, ids):
        import torch
        unique = torch.empty(ids.numel() + self.num_items() + 1, dtype=torch.int32, device=ids.device)
        m = self.r.lib.refcuda_ht_fill_duplicates(self.h, ids.data_ptr(), ids.numel(), unique.data_ptr())
        return unique[:m]

    def num_items(self):
        return int(self.r.lib.refcuda_ht_num_items(self.h))

    def map_edges(self, src, dst):
        import torch
        ns, nd = torch.empty_like(src), torch.empty_like(dst)
        self.r.lib.refcuda_map_edges(self.h, src.data_ptr(), ns.data_ptr(), dst.data_ptr(), nd.data_ptr(), src.numel())
        return ns, nd
