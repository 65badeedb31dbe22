"""ctypes binding of the kernel C-ABI (include/fgnn_kernels.h) for torch CUDA tensors.

PyTorch is plumbing here: it owns the device buffers and the stream; every
operation below is one call into libfgnn_kernels.so.  There is no fallback:
a missing library or a non-zero status raises.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "libfgnn_kernels.so")
EMPTY = 0xFFFFFFFF
CHAIN_WS_BYTES = 16 + 8 * 4096

SAMPLE_KHOP0, SAMPLE_KHOP1, SAMPLE_WEIGHTED, SAMPLE_RANDOM_WALK = 0, 1, 2, 3
SAMPLE_WEIGHTED_PREFIX, SAMPLE_KHOP2, SAMPLE_WEIGHTED_HASH_DEDUP = 4, 5, 6


class FgnnRng(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("batch_key", C.c_uint64), ("tag", C.c_uint32)]


MAX_LAYERS = 8
_P8 = C.c_void_p * MAX_LAYERS
_U8 = C.c_uint32 * MAX_LAYERS


class SamplePlan(C.Structure):
    """fgnn_sample_plan (include/fgnn_kernels.h)"""
    _fields_ = [("sample_type", C.c_int32), ("num_layers", C.c_uint32), ("fanout", _U8), ("in_max", _U8),
                ("indptr", C.c_void_p), ("indices", C.c_void_p), ("prob_table", C.c_void_p),
                ("alias_table", C.c_void_p), ("prob_prefix_table", C.c_void_p), ("walk_len", C.c_uint32),
                ("num_walk", C.c_uint32), ("restart_prob", C.c_double), ("seed", C.c_uint64),
                ("table", C.c_void_p), ("capacity", C.c_size_t), ("num_items", C.c_void_p),
                ("chain_ws", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("dst", _P8), ("pos", _P8), ("version", C.c_uint32), ("num_nodes", C.c_uint32),
                ("rank_ws", C.c_void_p)]


class SampleOut(C.Structure):
    """fgnn_sample_out (include/fgnn_kernels.h)"""
    _fields_ = [("n2o", C.c_void_p), ("row", _P8), ("col", _P8), ("data", _P8), ("counts", C.c_void_p)]


class CacheLayout(C.Structure):
    """fgnn_cache_layout (include/fgnn_kernels.h): replicated head + striped tail of the feature cache"""
    _fields_ = [("table", C.c_void_p), ("shards", C.c_void_p), ("num_shards", C.c_uint32), ("self_shard", C.c_uint32),
                ("replica", C.c_void_p), ("num_replicated", C.c_uint32), ("miss_src", C.c_void_p),
                ("miss_mask", C.c_uint64), ("row_bytes", C.c_size_t), ("defer_ws", C.c_void_p)]


class KernelError(RuntimeError):
    pass


_vp = C.c_void_p
_u32 = C.c_uint32
_u64 = C.c_uint64
_sz = C.c_size_t

_SIGS = {
    "fgnn_k_sample_khop": [C.c_int, _vp, _vp, _vp, _u32, _vp, _u32, FgnnRng, _vp, _vp, _vp, _vp, _vp, _vp],
    "fgnn_k_sample_replace": [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _u32, FgnnRng, _vp, _vp, _vp,
                              _vp, _vp, _sz, _vp, _vp],
    "fgnn_k_sample_replace_ranked": [C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _u32, _vp, _u32, FgnnRng, _vp, _vp, _vp,
                                     _vp, _vp, _sz, _vp, _vp, _sz, _vp],
    "fgnn_k_sample_weighted_hash_dedup": [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _u32, FgnnRng, _vp, _vp, _vp, _vp,
                                          _vp, _vp],
    "fgnn_k_sample_random_walk": [_vp, _vp, _vp, _u32, _vp, _u32, C.c_double, _u32, _u32, FgnnRng, _vp, _vp,
                                  _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp],
    "fgnn_k_ht_reset": [_vp, _sz, _vp, _vp],
    "fgnn_k_ht_fill_unique": [_vp, _sz, _vp, _u32, _vp, _vp, _vp, _vp],
    "fgnn_k_ht_fill_duplicates": [_vp, _sz, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
    "fgnn_k_ht_map": [_vp, _sz, _vp, _vp, _u32, _vp, _vp, _vp],
    "fgnn_k_ht_fill_duplicates_map": [_vp, _sz, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "fgnn_k_sample_batch": [C.POINTER(SamplePlan), C.POINTER(SampleOut), _vp, _u32, _vp, _u64, _vp],
    "fgnn_k_sample_batch_multi": [_vp, _vp, _vp, _vp, _vp, _u32, _vp],
    "fgnn_k_cache_table_build": [_vp, _sz, _vp, _sz, _vp],
    "fgnn_k_cache_split": [_vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "fgnn_k_row_copy": [_vp, _vp, _vp, _vp, _u64, _u32, _vp, _sz, _vp],
    "fgnn_k_gather_cached": [_vp, _vp, _u32, _vp, _vp, _vp, _u32, _vp, _u64, _sz, _vp, _vp],
    "fgnn_k_gather_cached_layout": [_vp, _vp, _u32, _vp, C.POINTER(CacheLayout), _vp, _vp, _vp],
    "fgnn_k_freq_count": [_vp, _vp, _u32, _vp, _vp],
    "fgnn_k_presc_rank": [_vp, _sz, _vp, _vp, _sz, _vp],
    "fgnn_k_shuffle": [_vp, _sz, _u64, _u64, _vp, _vp, _sz, _vp],
    "fgnn_k_build_alias_table": [_vp, _vp, _sz, _sz, _vp, _vp, _vp, _vp, _sz, _vp],
    "fgnn_k_build_prefix_table": [_vp, _sz, _vp, _vp, _vp],
    "fgnn_k_out_degree": [_vp, _sz, _vp, _sz, _vp],
    "fgnn_k_rank_by_degree": [_vp, _sz, _sz, _vp, _vp, _vp, _sz, _vp],
    "fgnn_k_rank_random": [_sz, _u64, _vp, _vp, _sz, _vp],
    "fgnn_k_row_len_sum": [_vp, _vp, _sz, _vp, _vp],
    "fgnn_k_rank_by_heuristic": [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp, _sz, _vp],
    "fgnn_k_trace_enable": [C.c_int, _sz],
    "fgnn_k_coo_to_csc": [_vp, _vp, _u32, _vp, _u32, C.c_int, _vp, _vp, _vp, _vp, _sz, _vp],
    "fgnn_k_shard_alloc": [C.POINTER(_vp), _sz],
    "fgnn_k_shard_free": [_vp],
    "fgnn_k_ipc_export": [_vp, _vp],
    "fgnn_k_ipc_open": [_vp, C.POINTER(_vp)],
    "fgnn_k_ipc_close": [_vp],
    "fgnn_k_enable_peer": [C.c_int],
}
_SIZE_FNS = {
    "fgnn_k_ht_capacity": [_sz],
    "fgnn_k_ht_bytes": [_sz],
    "fgnn_k_sample_replace_workspace_bytes": [_u32, _u32],
    "fgnn_k_seed_rank_workspace_bytes": [_sz],
    "fgnn_k_gather_defer_workspace_bytes": [_u32],
    "fgnn_k_sample_random_walk_workspace_bytes": [_u32, _u32],
    "fgnn_k_presc_rank_workspace_bytes": [_sz],
    "fgnn_k_shuffle_workspace_bytes": [_sz],
    "fgnn_k_alias_table_workspace_bytes": [_sz],
    "fgnn_k_rank_random_workspace_bytes": [_sz],
    "fgnn_k_coo_to_csc_workspace_bytes": [_u32, _u32],
    "fgnn_k_rank_heuristic_workspace_bytes": [_sz, _sz, _sz],
}

_lib = None


def exported_symbols():
    return list(_SIGS) + list(_SIZE_FNS) + ["fgnn_k_version", "fgnn_k_error_string", "fgnn_k_launch_count",
                                            "fgnn_k_trace_dump", "fgnn_k_ht_next_version"]


def load(path=None):
    """Load libfgnn_kernels.so (raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("FGNN_KERNELS_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise KernelError("%s not found: run `python fgnn-artifacts_b200/build.py` (no CPU fallback exists)" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    for name, args in _SIZE_FNS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _sz
    lib.fgnn_k_version.restype = C.c_char_p
    lib.fgnn_k_error_string.restype = C.c_char_p
    lib.fgnn_k_error_string.argtypes = [C.c_int]
    lib.fgnn_k_launch_count.restype = _u64
    lib.fgnn_k_trace_dump.restype = C.c_long
    lib.fgnn_k_trace_dump.argtypes = [_sz, _vp, _vp, _vp]
    _lib = lib
    return lib


def _check(code, what):
    if code != 0:
        raise KernelError("%s failed: %s (%d)" % (what, load().fgnn_k_error_string(code).decode(), code))


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, int):
        return t
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(load().fgnn_k_launch_count())


def new_chain_ws(device="cuda"):
    return torch.zeros(CHAIN_WS_BYTES // 4, dtype=torch.int32, device=device)


def rng(seed, batch_key, tag):
    return FgnnRng(seed & 0xFFFFFFFFFFFFFFFF, batch_key & 0xFFFFFFFFFFFFFFFF, tag & 0xFFFFFFFF)


# ---------------------------------------------------------------------------
# thin wrappers (tensors are int32-typed views of uint32 data)
# ---------------------------------------------------------------------------
def sample_khop(variant, indptr, indices, inp, n_max, d_n, fanout, r, out_src, out_dst, out_src_local,
                d_num_out, chain_ws):
    _check(load().fgnn_k_sample_khop(variant, _ptr(indptr), _ptr(indices), _ptr(inp), n_max, _ptr(d_n), fanout,
                                     r, _ptr(out_src), _ptr(out_dst), _ptr(out_src_local), _ptr(d_num_out),
                                     _ptr(chain_ws), _stream()), "sample_khop")


def sample_replace_workspace_bytes(n_max, fanout):
    return int(load().fgnn_k_sample_replace_workspace_bytes(n_max, fanout))


def seed_rank_workspace_bytes(num_nodes):
    return int(load().fgnn_k_seed_rank_workspace_bytes(num_nodes))


def new_rank_ws(num_nodes, device="cuda"):
    """Zero-initialised workspace of the rank-by-bitmap seed ordering (None when the id space is too large)."""
    nb = seed_rank_workspace_bytes(num_nodes)
    return torch.zeros(nb, dtype=torch.uint8, device=device) if nb else None


def sample_replace(kind, indptr, indices, prob, alias, prefix, inp, n_max, d_n, fanout, r, out_src, out_dst,
                   out_src_local, d_num_out, workspace, chain_ws, rank_ws=None, num_nodes=0):
    """khop1 / weighted alias / prefix sampler; with `rank_ws` (new_rank_ws) the seeds are ordered by the
    rank-by-bitmap kernels instead of a radix sort."""
    _check(load().fgnn_k_sample_replace_ranked(kind, _ptr(indptr), _ptr(indices), _ptr(prob), _ptr(alias), _ptr(prefix),
                                               _ptr(inp), n_max, _ptr(d_n), fanout, r, _ptr(out_src), _ptr(out_dst),
                                               _ptr(out_src_local), _ptr(d_num_out), _ptr(workspace),
                                               workspace.numel() * workspace.element_size(), _ptr(chain_ws),
                                               _ptr(rank_ws), num_nodes if rank_ws is not None else 0, _stream()),
           "sample_replace")


def sample_weighted_hash_dedup(indptr, indices, prob, alias, inp, n_max, d_n, fanout, r, out_src, out_dst,
                               out_src_local, d_num_out, chain_ws):
    _check(load().fgnn_k_sample_weighted_hash_dedup(_ptr(indptr), _ptr(indices), _ptr(prob), _ptr(alias),
                                                    _ptr(inp), n_max, _ptr(d_n), fanout, r, _ptr(out_src),
                                                    _ptr(out_dst), _ptr(out_src_local), _ptr(d_num_out),
                                                    _ptr(chain_ws), _stream()), "sample_weighted_hash_dedup")


def sample_random_walk_workspace_bytes(n_max, K):
    return int(load().fgnn_k_sample_random_walk_workspace_bytes(n_max, K))


def sample_random_walk(indptr, indices, inp, n_max, d_n, walk_len, restart_prob, num_walk, K, r, out_src,
                       out_dst, out_src_local, out_data, d_num_out, tmp_src, tmp_dst, workspace, chain_ws):
    _check(load().fgnn_k_sample_random_walk(_ptr(indptr), _ptr(indices), _ptr(inp), n_max, _ptr(d_n), walk_len,
                                            float(restart_prob), num_walk, K, r, _ptr(out_src), _ptr(out_dst),
                                            _ptr(out_src_local), _ptr(out_data), _ptr(d_num_out), _ptr(tmp_src),
                                            _ptr(tmp_dst), _ptr(workspace),
                                            workspace.numel() * workspace.element_size(), _ptr(chain_ws),
                                            _stream()), "sample_random_walk")


def ht_capacity(max_items):
    return int(load().fgnn_k_ht_capacity(max_items))


def ht_bytes(capacity):
    return int(load().fgnn_k_ht_bytes(capacity))


def ht_reset(table, capacity, d_num_items):
    _check(load().fgnn_k_ht_reset(_ptr(table), capacity, _ptr(d_num_items), _stream()), "ht_reset")


def ht_fill_unique(table, capacity, inp, n_max, d_n, n2o, d_num_items):
    _check(load().fgnn_k_ht_fill_unique(_ptr(table), capacity, _ptr(inp), n_max, _ptr(d_n), _ptr(n2o),
                                        _ptr(d_num_items), _stream()), "ht_fill_unique")


def ht_fill_duplicates(table, capacity, inp, n_max, d_n, pos, n2o, d_num_items, chain_ws):
    _check(load().fgnn_k_ht_fill_duplicates(_ptr(table), capacity, _ptr(inp), n_max, _ptr(d_n), _ptr(pos),
                                            _ptr(n2o), _ptr(d_num_items), _ptr(chain_ws), _stream()),
           "ht_fill_duplicates")


def ht_fill_duplicates_map(table, capacity, inp, n_max, d_n, pos, n2o, d_num_items, out_local, chain_ws):
    _check(load().fgnn_k_ht_fill_duplicates_map(_ptr(table), capacity, _ptr(inp), n_max, _ptr(d_n), _ptr(pos),
                                                _ptr(n2o), _ptr(d_num_items), _ptr(out_local), _ptr(chain_ws),
                                                _stream()), "ht_fill_duplicates_map")


def sample_batch(plan, out, seeds, n_seeds_max, d_n_seeds, batch_key):
    """DoGPUSample for one mini-batch (fgnn_k_sample_batch) on the current torch stream."""
    _check(load().fgnn_k_sample_batch(C.byref(plan), C.byref(out), _ptr(seeds), n_seeds_max, _ptr(d_n_seeds),
                                      batch_key & 0xFFFFFFFFFFFFFFFF, _stream()), "sample_batch")


MAX_SUPER = 8


def sample_batch_multi(plans, outs, seeds, n_seeds, batch_keys):
    """DoGPUSample for up to MAX_SUPER mini-batches of one configuration in one call (fgnn_k_sample_batch_multi):
    for the uniform k-hop sampler every layer is two launches for all of them."""
    k = len(plans)
    assert 0 < k <= MAX_SUPER and len(outs) == len(seeds) == len(n_seeds) == len(batch_keys) == k
    pp = (C.c_void_p * k)(*[C.addressof(p) for p in plans])
    oo = (C.c_void_p * k)(*[C.addressof(o) for o in outs])
    ss = (C.c_void_p * k)(*[_ptr(s) for s in seeds])
    nn = (C.c_uint32 * k)(*n_seeds)
    kk = (C.c_uint64 * k)(*[b & 0xFFFFFFFFFFFFFFFF for b in batch_keys])
    _check(load().fgnn_k_sample_batch_multi(pp, oo, ss, nn, kk, k, _stream()), "sample_batch_multi")


def ht_next_version(state, table, capacity):
    """Next version tag (1..126) of a hash table; `state` is a ctypes c_uint32 kept by the table's owner.  The
    table is cleared on the current stream when the tags wrap."""
    lib = load()
    lib.fgnn_k_ht_next_version.argtypes = [C.POINTER(C.c_uint32), _vp, _sz, _vp]
    lib.fgnn_k_ht_next_version.restype = C.c_uint32
    return int(lib.fgnn_k_ht_next_version(C.byref(state), _ptr(table), capacity, _stream()))


def ht_map(table, capacity, glob, pos, n_max, d_n, out_local):
    _check(load().fgnn_k_ht_map(_ptr(table), capacity, _ptr(glob), _ptr(pos), n_max, _ptr(d_n), _ptr(out_local),
                                _stream()), "ht_map")


def cache_table_build(table, num_nodes, rank, num_cached):
    _check(load().fgnn_k_cache_table_build(_ptr(table), num_nodes, _ptr(rank), num_cached, _stream()),
           "cache_table_build")


def cache_split(table, nodes, n_max, d_n, miss_src, miss_dst, cache_src, cache_dst, d_counts, chain_ws):
    _check(load().fgnn_k_cache_split(_ptr(table), _ptr(nodes), n_max, _ptr(d_n), _ptr(miss_src), _ptr(miss_dst),
                                     _ptr(cache_src), _ptr(cache_dst), _ptr(d_counts), _ptr(chain_ws), _stream()),
           "cache_split")


def row_copy(dst, dst_index, src, src_index, n_max, d_n, row_bytes, src_mask=0xFFFFFFFFFFFFFFFF):
    _check(load().fgnn_k_row_copy(_ptr(dst), _ptr(dst_index), _ptr(src), _ptr(src_index), src_mask, n_max,
                                  _ptr(d_n), row_bytes, _stream()), "row_copy")


def gather_cached(out, nodes, n_max, d_n, table, shards, num_shards, miss_src, row_bytes, d_stats=None,
                  miss_mask=0xFFFFFFFFFFFFFFFF):
    _check(load().fgnn_k_gather_cached(_ptr(out), _ptr(nodes), n_max, _ptr(d_n), _ptr(table), _ptr(shards),
                                       num_shards, _ptr(miss_src), miss_mask, row_bytes, _ptr(d_stats), _stream()),
           "gather_cached")


def gather_cached_layout(out, nodes, n_max, d_n, layout, d_stats=None, d_remote=None):
    """The cache-aware gather over a hybrid layout (hottest slots replicated, the rest striped over the shards)."""
    _check(load().fgnn_k_gather_cached_layout(_ptr(out), _ptr(nodes), n_max, _ptr(d_n), C.byref(layout),
                                              _ptr(d_stats), _ptr(d_remote), _stream()), "gather_cached_layout")


def freq_count(freq, nodes, n_max, d_n):
    _check(load().fgnn_k_freq_count(_ptr(freq), _ptr(nodes), n_max, _ptr(d_n), _stream()), "freq_count")


def presc_rank_workspace_bytes(num_nodes):
    return int(load().fgnn_k_presc_rank_workspace_bytes(num_nodes))


def presc_rank(freq, num_nodes, rank, workspace):
    _check(load().fgnn_k_presc_rank(_ptr(freq), num_nodes, _ptr(rank), _ptr(workspace),
                                    workspace.numel() * workspace.element_size(), _stream()), "presc_rank")


def build_alias_table(indptr, indices, num_nodes, num_edges, weights, prob, alias, workspace=None):
    """prob_table / alias_table from per-edge weights (create_alias_table.cc:96-180), on the GPU."""
    import torch
    if workspace is None:
        workspace = torch.empty(int(load().fgnn_k_alias_table_workspace_bytes(num_edges)), dtype=torch.uint8,
                                device=prob.device)
    _check(load().fgnn_k_build_alias_table(_ptr(indptr), _ptr(indices), num_nodes, num_edges, _ptr(weights),
                                           _ptr(prob), _ptr(alias), _ptr(workspace), workspace.numel(), _stream()),
           "build_alias_table")


def build_prefix_table(indptr, num_nodes, weights, prefix):
    _check(load().fgnn_k_build_prefix_table(_ptr(indptr), num_nodes, _ptr(weights), _ptr(prefix), _stream()),
           "build_prefix_table")


def out_degree(indices, num_edges, out_deg, num_nodes):
    _check(load().fgnn_k_out_degree(_ptr(indices), num_edges, _ptr(out_deg), num_nodes, _stream()), "out_degree")


def rank_by_degree(indices, num_edges, num_nodes, out_deg, rank, workspace):
    _check(load().fgnn_k_rank_by_degree(_ptr(indices), num_edges, num_nodes, _ptr(out_deg), _ptr(rank),
                                        _ptr(workspace), workspace.numel(), _stream()), "rank_by_degree")


def rank_random(num_nodes, seed, rank, workspace=None):
    import torch
    if workspace is None:
        workspace = torch.empty(int(load().fgnn_k_rank_random_workspace_bytes(num_nodes)), dtype=torch.uint8,
                                device=rank.device)
    _check(load().fgnn_k_rank_random(num_nodes, seed & 0xFFFFFFFFFFFFFFFF, _ptr(rank), _ptr(workspace),
                                     workspace.numel(), _stream()), "rank_random")


def rank_by_heuristic(indptr, indices, num_nodes, num_edges, train_set, num_train, rank):
    """cache_by_heuristic ranking on the GPU; one host read-back (the neighbour count) sizes the workspace."""
    import torch
    total = torch.zeros(1, dtype=torch.int64, device=rank.device)
    _check(load().fgnn_k_row_len_sum(_ptr(indptr), _ptr(train_set), num_train, _ptr(total), _stream()), "row_len_sum")
    n_nbr = int(total.item())
    ws = torch.empty(int(load().fgnn_k_rank_heuristic_workspace_bytes(num_nodes, num_train, n_nbr)),
                     dtype=torch.uint8, device=rank.device)
    _check(load().fgnn_k_rank_by_heuristic(_ptr(indptr), _ptr(indices), num_nodes, num_edges, _ptr(train_set),
                                           num_train, n_nbr, _ptr(rank), _ptr(ws), ws.numel(), _stream()),
           "rank_by_heuristic")
    return n_nbr


TRACE_LABELS = {0: "batch_begin", 1: "table_reset", 2: "fill_seeds", 900: "gather_begin", 901: "gather_end"}
TRACE_LAYER_OPS = {1: "sample", 2: "insert", 3: "compact", 4: "map"}


def trace_enable(on=True, max_records=8192):
    """Record a CUDA event after every launch of fgnn_k_sample_batch / fgnn_k_gather_cached (profiling aid)."""
    _check(load().fgnn_k_trace_enable(int(bool(on)), max_records), "trace_enable")


def trace_dump(max_records=8192):
    """-> [(name, stream handle, ms since trace_enable)] in enqueue order; a record marks the END of the launch."""
    labels = (C.c_int * max_records)()
    streams = (C.c_uint64 * max_records)()
    ms = (C.c_float * max_records)()
    n = int(load().fgnn_k_trace_dump(max_records, labels, streams, ms))
    out = []
    for i in range(n):
        lab = labels[i]
        name = TRACE_LABELS.get(lab) or "L%d_%s" % (lab // 100 - 1, TRACE_LAYER_OPS.get(lab % 100, "?"))
        out.append((name, int(streams[i]), float(ms[i])))
    return out


def coo_to_csc(row, col, e_max, d_e, num_dst, col_sorted, indptr, indices=None, edge_ids=None, workspace=None):
    """(row, col) of one sampled layer -> CSC indptr / indices / edge_ids (include/fgnn_kernels.h)."""
    import torch
    if not col_sorted and workspace is None and e_max:
        workspace = torch.empty(int(load().fgnn_k_coo_to_csc_workspace_bytes(e_max, num_dst)), dtype=torch.uint8,
                                device=indptr.device)
    _check(load().fgnn_k_coo_to_csc(_ptr(row), _ptr(col), e_max, _ptr(d_e), num_dst, int(bool(col_sorted)),
                                    _ptr(indptr), _ptr(indices), _ptr(edge_ids), _ptr(workspace),
                                    workspace.numel() if workspace is not None else 0, _stream()), "coo_to_csc")


def shuffle_workspace_bytes(n):
    return int(load().fgnn_k_shuffle_workspace_bytes(n))


def shuffle(train_set, n, seed, epoch, out, workspace):
    _check(load().fgnn_k_shuffle(_ptr(train_set), n, seed & 0xFFFFFFFFFFFFFFFF, epoch, _ptr(out), _ptr(workspace),
                                 workspace.numel() * workspace.element_size(), _stream()), "shuffle")


def shard_alloc(nbytes):
    p = _vp()
    _check(load().fgnn_k_shard_alloc(C.byref(p), nbytes), "shard_alloc")
    return p.value


def shard_free(ptr):
    _check(load().fgnn_k_shard_free(ptr), "shard_free")


def ipc_export(ptr):
    buf = C.create_string_buffer(64)
    _check(load().fgnn_k_ipc_export(ptr, buf), "ipc_export")
    return buf.raw


def ipc_open(handle_bytes):
    p = _vp()
    buf = C.create_string_buffer(handle_bytes, 64)
    _check(load().fgnn_k_ipc_open(buf, C.byref(p)), "ipc_open")
    return p.value


def ipc_close(ptr):
    _check(load().fgnn_k_ipc_close(ptr), "ipc_close")


def enable_peer(peer_device):
    _check(load().fgnn_k_enable_peer(peer_device), "enable_peer")
