"""Device-resident driver of one mini-batch of the hot path, on top of the kernel
C-ABI.  Mirrors DoGPUSample + DoGetCacheMissIndex + DoCacheFeatureCopy
(samgraph/common/cuda/cuda_loops.cc:50-267, dist_loops.cc:271-323,713-846) but
enqueues the whole batch without a host round trip: every count stays in HBM
(n_max / d_n convention of include/fgnn_kernels.h) and buffers are sized once
from PredictNumNodes (common.cc:330-339).

Used by bench.py (kernel-resident leg) and the GPU tests.  The production host
runtime is the C++ engine in csrc/runtime (samgraph_* C-ABI); this class is the
same sequence of kernel calls driven from Python.
"""
import torch

from . import kernels as K

SAMPLE_TYPES = {"khop0": 0, "khop1": 1, "weighted_khop": 2, "random_walk": 3, "weighted_khop_prefix": 4,
                "khop2": 5, "weighted_khop_hash_dedup": 6}


def predict_num_nodes(batch, fanouts, upto=None):
    upto = len(fanouts) if upto is None else upto
    count = batch
    for i in range(upto - 1, -1, -1):
        count += count * fanouts[i]
    return count


class HotPath:
    def __init__(self, indptr, indices, num_nodes, fanouts, batch_size, sample_type="khop2", seed=0x5EED,
                 prob_table=None, alias_table=None, prefix_table=None, rw=None, device="cuda"):
        K.load()
        self.dev = device
        self.indptr, self.indices = indptr, indices
        self.num_nodes = num_nodes
        self.fanouts = list(fanouts)
        self.L = len(self.fanouts)
        self.batch_size = batch_size
        self.sample_type = sample_type
        self.seed = seed
        self.prob, self.alias, self.prefix = prob_table, alias_table, prefix_table
        self.rw = rw or {}
        i32 = dict(dtype=torch.int32, device=device)
        # layer input bounds: S_l for l = L-1 .. 0 (cuda_loops.cc:87)
        self.in_max = [0] * self.L
        cur = batch_size
        for i in range(self.L - 1, -1, -1):
            self.in_max[i] = cur
            cur = cur * (self.fanouts[i] + 1)
        self.max_nodes = predict_num_nodes(batch_size, self.fanouts)
        self.edge_max = [self.in_max[i] * self.fanouts[i] for i in range(self.L)]
        self.cap = K.ht_capacity(self.max_nodes)
        self.table = torch.empty(K.ht_bytes(self.cap) // 4, **i32)
        self.n2o = torch.empty(self.max_nodes + 1, **i32)       # == input_nodes after the last layer
        self.num_items = torch.zeros(1, **i32)
        self.chain = K.new_chain_ws(device)
        self.dst = [torch.empty(max(1, m), **i32) for m in self.edge_max]
        self.col = [torch.empty(max(1, m), **i32) for m in self.edge_max]
        self.row = [torch.empty(max(1, m), **i32) for m in self.edge_max]
        self.pos = [torch.empty(max(1, m), **i32) for m in self.edge_max]
        self.data = [torch.empty(max(1, m), **i32) for m in self.edge_max] if sample_type == "random_walk" else None
        # counts[l] = (num_dst, num_edge, num_src) of layer l
        self.counts = torch.zeros((self.L, 3), **i32)
        self.ws = None
        if sample_type in ("khop1", "weighted_khop", "weighted_khop_prefix"):
            nb = max(K.sample_replace_workspace_bytes(self.in_max[i], self.fanouts[i]) for i in range(self.L))
            self.ws = torch.empty(nb, dtype=torch.uint8, device=device)
        elif sample_type == "random_walk":
            nb = max(K.sample_random_walk_workspace_bytes(self.in_max[i], self.fanouts[i]) for i in range(self.L))
            self.ws = torch.empty(nb, dtype=torch.uint8, device=device)
        # cache state (set by build_cache)
        self.cache_table = None
        self.shards = None
        self.shard_ptrs = None
        self.num_shards = 1
        self.miss_src = None
        self.miss_mask = 0xFFFFFFFFFFFFFFFF
        self.row_bytes = 0
        self.feat_out = None
        self.label_src = None
        self.label_out = None
        self.stats = torch.zeros(2, dtype=torch.int64, device=device)
        self.split_counts = torch.zeros(2, **i32)

    # ------------------------------------------------------------------
    def sample(self, seeds, n_seeds, batch_key):
        """DoGPUSample: seeds (device int32/u32) -> per-layer (row, col[, data]) + input_nodes."""
        K.ht_reset(self.table, self.cap, self.num_items)
        K.ht_fill_unique(self.table, self.cap, seeds, n_seeds, None, self.n2o, self.num_items)
        st = SAMPLE_TYPES[self.sample_type]
        for i in range(self.L - 1, -1, -1):
            f = self.fanouts[i]
            n_in = self.counts[i, 0:1]
            n_edge = self.counts[i, 1:2]
            n_in.copy_(self.num_items)                      # layer input = running unique list
            r = K.rng(self.seed, batch_key, i)
            if st in (0, 5):
                K.sample_khop(0 if st == 0 else 2, self.indptr, self.indices, self.n2o, self.in_max[i], n_in, f, r,
                              None, self.dst[i], self.col[i], n_edge, self.chain)
            elif st in (1, 2, 4):
                K.sample_replace(st, self.indptr, self.indices, self.prob, self.alias, self.prefix, self.n2o,
                                 self.in_max[i], n_in, f, r, None, self.dst[i], self.col[i], n_edge, self.ws,
                                 self.chain)
            elif st == 6:
                K.sample_weighted_hash_dedup(self.indptr, self.indices, self.prob, self.alias, self.n2o,
                                             self.in_max[i], n_in, f, r, None, self.dst[i], self.col[i], n_edge,
                                             self.chain)
            else:
                K.sample_random_walk(self.indptr, self.indices, self.n2o, self.in_max[i], n_in,
                                     self.rw["random_walk_length"], self.rw["random_walk_restart_prob"],
                                     self.rw["num_random_walk"], f, r, None, self.dst[i], self.col[i], self.data[i],
                                     n_edge, None, None, self.ws, self.chain)
            K.ht_fill_duplicates(self.table, self.cap, self.dst[i], self.edge_max[i], n_edge, self.pos[i], self.n2o,
                                 self.num_items, self.chain)
            K.ht_map(self.table, self.cap, None, self.pos[i], self.edge_max[i], n_edge, self.row[i])
            self.counts[i, 2:3].copy_(self.num_items)

    # ------------------------------------------------------------------
    def presample_count(self, freq):
        K.freq_count(freq, self.n2o, self.max_nodes, self.num_items)

    def build_cache(self, ranking_nodes, cache_percentage, feat_src, row_bytes, feat_mask=0xFFFFFFFFFFFFFFFF,
                    num_shards=1, shard_id=0, peer_ptrs=None):
        """GPUCacheManager / DistCacheManager ctor (cuda_cache_manager_host.cc:60-127): node->slot table and
        the cached rows.  With num_shards > 1 only rows slot % num_shards == shard_id are stored locally."""
        V = self.num_nodes
        self.row_bytes = row_bytes
        self.num_cached = int(V * cache_percentage)
        self.cache_table = torch.empty(V, dtype=torch.int32, device=self.dev)
        K.cache_table_build(self.cache_table, V, ranking_nodes, self.num_cached)
        self.num_shards = num_shards
        local_rows = (self.num_cached - shard_id + num_shards - 1) // num_shards if self.num_cached > shard_id else 0
        self.cache = torch.empty((max(1, local_rows), row_bytes), dtype=torch.uint8, device=self.dev)
        if local_rows:
            idx = ranking_nodes[shard_id:self.num_cached:num_shards].contiguous()
            K.row_copy(self.cache, None, feat_src, idx, local_rows, None, row_bytes, feat_mask)
        self.miss_src, self.miss_mask = feat_src, feat_mask
        if peer_ptrs is None:
            peer_ptrs = [self.cache.data_ptr()]
        self.shard_ptrs = torch.tensor(peer_ptrs, dtype=torch.int64, device=self.dev)
        self.feat_out = torch.empty((self.max_nodes, row_bytes), dtype=torch.uint8, device=self.dev)
        return local_rows

    def set_labels(self, label_src):
        self.label_src = label_src
        self.label_out = torch.empty(self.batch_size, dtype=torch.int64, device=self.dev)

    def extract(self, seeds=None, n_seeds=0):
        """DoCacheFeatureCopy + DoCPULabelExtractAndCopy, fused gather."""
        K.gather_cached(self.feat_out, self.n2o, self.max_nodes, self.num_items, self.cache_table, self.shard_ptrs,
                        self.num_shards, self.miss_src, self.row_bytes, self.stats, self.miss_mask)
        if self.label_src is not None and seeds is not None:
            K.row_copy(self.label_out, None, self.label_src, seeds, n_seeds, None, 8)

    def split(self, bufs):
        """GetMissCacheIndex: explicit index lists (reference-compatible output)."""
        K.cache_split(self.cache_table, self.n2o, self.max_nodes, self.num_items, bufs[0], bufs[1], bufs[2], bufs[3],
                      self.split_counts, self.chain)

    def step(self, seeds, n_seeds, batch_key):
        self.sample(seeds, n_seeds, batch_key)
        self.extract(seeds, n_seeds)
