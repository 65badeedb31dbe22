"""Device-resident driver of one mini-batch of the hot path, on top of the kernel
C-ABI.  Mirrors DoGPUSample + DoGetCacheMissIndex + DoCacheFeatureCopy
(samgraph/common/cuda/cuda_loops.cc:50-267, dist_loops.cc:271-323,713-846) but
enqueues the whole batch without a host round trip: every count stays in HBM
(n_max / d_n convention of include/fgnn_kernels.h) and buffers are sized once
from PredictNumNodes (common.cc:330-339).

Used by bench.py (kernel-resident leg) and the GPU tests.  The production host
runtime is the C++ engine in csrc/runtime (samgraph_* C-ABI); this class issues
the same two calls per batch (fgnn_k_sample_batch, fgnn_k_gather_cached) from
Python.  Like the engine's sampler it owns `num_slots` independent sets of
scratch + output buffers, so several batches can be in flight on different
streams.
"""
import ctypes as C

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


class Slot:
    """Scratch + outputs of one batch in flight (SampleSlot + TaskBlock of csrc/runtime/rt_engine.cc)."""

    def __init__(self, hp):
        i32 = dict(dtype=torch.int32, device=hp.dev)
        L = hp.L
        self.table = torch.empty(K.ht_bytes(hp.cap) // 4, **i32)
        self.n2o = torch.empty(hp.max_nodes + 1, **i32)       # == input_nodes after the last layer
        self.num_items = torch.zeros(1, **i32)
        self.chain = K.new_chain_ws(hp.dev)
        self.dst = [torch.empty(max(1, m), **i32) for m in hp.edge_max]
        self.col = [torch.empty(max(1, m), **i32) for m in hp.edge_max]
        self.row = [torch.empty(max(1, m), **i32) for m in hp.edge_max]
        self.pos = [torch.empty(max(1, m), **i32) for m in hp.edge_max]
        self.data = [torch.empty(max(1, m), **i32) for m in hp.edge_max] if hp.sample_type == "random_walk" else None
        # counts[l] = (num_dst, num_edge, num_src) of layer l
        self.counts = torch.zeros((L, 3), **i32)
        self.ws = None
        self.rank_ws = None
        if hp.sample_type in ("khop1", "weighted_khop", "weighted_khop_prefix"):
            nb = max(K.sample_replace_workspace_bytes(hp.in_max[i], hp.fanouts[i]) for i in range(L))
            self.ws = torch.empty(nb, dtype=torch.uint8, device=hp.dev)
            if hp.rank_by_bitmap:
                self.rank_ws = K.new_rank_ws(hp.num_nodes, hp.dev)
        elif hp.sample_type == "random_walk":
            nb = max(K.sample_random_walk_workspace_bytes(hp.in_max[i], hp.fanouts[i]) for i in range(L))
            self.ws = torch.empty(nb, dtype=torch.uint8, device=hp.dev)
        pl = K.SamplePlan()
        pl.sample_type = SAMPLE_TYPES[hp.sample_type]
        pl.num_layers = L
        for i in range(L):
            pl.fanout[i] = hp.fanouts[i]
            pl.in_max[i] = hp.in_max[i]
            pl.dst[i] = self.dst[i].data_ptr()
            pl.pos[i] = self.pos[i].data_ptr()
        pl.indptr, pl.indices = hp.indptr.data_ptr(), hp.indices.data_ptr()
        pl.prob_table = hp.prob.data_ptr() if hp.prob is not None else None
        pl.alias_table = hp.alias.data_ptr() if hp.alias is not None else None
        pl.prob_prefix_table = hp.prefix.data_ptr() if hp.prefix is not None else None
        pl.walk_len = hp.rw.get("random_walk_length", 0)
        pl.num_walk = hp.rw.get("num_random_walk", 0)
        pl.restart_prob = hp.rw.get("random_walk_restart_prob", 0.0)
        pl.seed = hp.seed & 0xFFFFFFFFFFFFFFFF
        pl.table, pl.capacity = self.table.data_ptr(), hp.cap
        pl.num_items, pl.chain_ws = self.num_items.data_ptr(), self.chain.data_ptr()
        pl.workspace = self.ws.data_ptr() if self.ws is not None else None
        pl.workspace_bytes = self.ws.numel() if self.ws is not None else 0
        pl.num_nodes = hp.num_nodes if self.rank_ws is not None else 0
        pl.rank_ws = self.rank_ws.data_ptr() if self.rank_ws is not None else None
        out = K.SampleOut()
        out.n2o, out.counts = self.n2o.data_ptr(), self.counts.data_ptr()
        for i in range(L):
            out.row[i], out.col[i] = self.row[i].data_ptr(), self.col[i].data_ptr()
            out.data[i] = self.data[i].data_ptr() if self.data is not None else None
        self.plan, self.out = pl, out
        self.ver_state = C.c_uint32(0)      # versioned table reset: fgnn_k_ht_next_version hands out the tags


class HotPath:
    def __init__(self, indptr, indices, num_nodes, fanouts, batch_size, sample_type="khop2", seed=0x5EED,
                 prob_table=None, alias_table=None, prefix_table=None, rw=None, device="cuda", num_slots=1,
                 ht_capacity=None):
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
        # layer input bounds: S_l for l = L-1 .. 0 (cuda_loops.cc:87)
        self.in_max = [0] * self.L
        cur = batch_size
        for i in range(self.L - 1, -1, -1):
            self.in_max[i] = cur
            cur = cur * (self.fanouts[i] + 1)
        self.max_nodes = predict_num_nodes(batch_size, self.fanouts)
        self.edge_max = [self.in_max[i] * self.fanouts[i] for i in range(self.L)]
        # ht_capacity: power-of-two override for experiments; must exceed the batch's unique-node count
        self.cap = ht_capacity if ht_capacity else K.ht_capacity(self.max_nodes)
        import os
        self.versioned = os.environ.get("FGNN_HT_VERSIONED", "1") != "0"
        self.rank_by_bitmap = os.environ.get("FGNN_SEED_RANK", "1") != "0"     # 0: order the seeds with the CUB sort
        self.slots = [Slot(self) for _ in range(max(1, num_slots))]
        self._alias_slot(0)
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
        self.split_counts = torch.zeros(2, dtype=torch.int32, device=device)

    def _alias_slot(self, s):
        """Single-slot callers (tests, smoke) read the results as attributes of the HotPath itself."""
        sl = self.slots[s]
        self.table, self.n2o, self.num_items, self.chain = sl.table, sl.n2o, sl.num_items, sl.chain
        self.dst, self.col, self.row, self.pos, self.data, self.counts = sl.dst, sl.col, sl.row, sl.pos, sl.data, sl.counts

    # ------------------------------------------------------------------
    def sample(self, seeds, n_seeds, batch_key, slot=0):
        """DoGPUSample: seeds (device int32/u32) -> per-layer (row, col[, data]) + input_nodes; one C call."""
        sl = self.slots[slot]
        self._next_version(sl)
        K.sample_batch(sl.plan, sl.out, seeds, n_seeds, None, batch_key)

    def _next_version(self, sl):
        sl.plan.version = K.ht_next_version(sl.ver_state, sl.table, self.cap) if self.versioned else 0

    def sample_multi(self, batches):
        """Super-batch: `batches` = [(seeds, n_seeds, batch_key, slot), ...] (distinct slots) enqueued together on
        the current stream with ONE call; results land in the slots exactly as `sample` would leave them."""
        slots = [self.slots[b[3]] for b in batches]
        assert len(set(b[3] for b in batches)) == len(batches)
        for sl in slots:
            self._next_version(sl)
        K.sample_batch_multi([sl.plan for sl in slots], [sl.out for sl in slots], [b[0] for b in batches],
                             [b[1] for b in batches], [b[2] for b in batches])

    # ------------------------------------------------------------------
    def presample_count(self, freq, slot=0):
        sl = self.slots[slot]
        K.freq_count(freq, sl.n2o, self.max_nodes, sl.num_items)

    def build_cache(self, ranking_nodes, cache_percentage, feat_src, row_bytes, feat_mask=0xFFFFFFFFFFFFFFFF,
                    num_shards=1, shard_id=0, peer_ptrs=None, fill_local=True, num_replicated=0, replica_ptr=None):
        """GPUCacheManager / DistCacheManager ctor (cuda_cache_manager_host.cc:60-127): node->slot table and
        the cached rows.  With num_shards > 1 only rows slot % num_shards == shard_id are stored locally;
        `peer_ptrs` then lists the base pointer of every shard (own + NVLink peer mappings, see
        fgnn_b200/partition.py) and, with fill_local=False, the caller has already filled its own shard.
        Hybrid layout: the hottest `num_replicated` slots live in `replica_ptr` on every GPU, only the slots
        behind them are striped (fgnn_cache_layout)."""
        V = self.num_nodes
        self.row_bytes = row_bytes
        self.num_cached = int(V * cache_percentage)
        self.cache_table = torch.empty(V, dtype=torch.int32, device=self.dev)
        K.cache_table_build(self.cache_table, V, ranking_nodes, self.num_cached)
        self.num_shards = num_shards
        self.shard_id = shard_id
        self.num_replicated = min(int(num_replicated), self.num_cached)
        self.replica_ptr = replica_ptr
        assert self.num_replicated == 0 or (replica_ptr and not fill_local)
        striped = self.num_cached - self.num_replicated
        local_rows = (striped - shard_id + num_shards - 1) // num_shards if striped > shard_id else 0
        self.cache = None
        if fill_local:
            self.cache = torch.empty((max(1, local_rows), row_bytes), dtype=torch.uint8, device=self.dev)
            if local_rows:
                idx = ranking_nodes[shard_id:self.num_cached:num_shards].contiguous()
                K.row_copy(self.cache, None, feat_src, idx, local_rows, None, row_bytes, feat_mask)
        self.miss_src, self.miss_mask = feat_src, feat_mask
        if peer_ptrs is None:
            assert fill_local and num_shards == 1
            peer_ptrs = [self.cache.data_ptr()]
        assert len(peer_ptrs) == num_shards
        self.shard_ptrs = torch.tensor(peer_ptrs, dtype=torch.int64, device=self.dev)
        lay = K.CacheLayout()
        lay.table, lay.shards = self.cache_table.data_ptr(), self.shard_ptrs.data_ptr()
        lay.num_shards, lay.self_shard = num_shards, shard_id
        lay.replica, lay.num_replicated = replica_ptr, self.num_replicated
        lay.miss_src, lay.miss_mask, lay.row_bytes = K._ptr(feat_src), feat_mask, row_bytes
        # peer rows in a second pass (fgnn_cache_layout.defer_ws): needed once several GPUs gather at the same time
        self.defer_ws = None
        if num_shards > 1:
            nb = int(K.load().fgnn_k_gather_defer_workspace_bytes(self.max_nodes))
            self.defer_ws = torch.zeros(nb, dtype=torch.uint8, device=self.dev)
            lay.defer_ws = self.defer_ws.data_ptr()
        self.layout = lay
        self.remote = torch.zeros(1, dtype=torch.int64, device=self.dev)
        if self.feat_out is None:
            self.feat_out = torch.empty((self.max_nodes, row_bytes), dtype=torch.uint8, device=self.dev)
        return local_rows

    def set_labels(self, label_src):
        self.label_src = label_src
        self.label_out = torch.empty(self.batch_size, dtype=torch.int64, device=self.dev)

    def gather(self, slot=0):
        """DoCacheFeatureCopy: the fused cache-aware feature gather of the slot's input_nodes."""
        sl = self.slots[slot]
        K.gather_cached_layout(self.feat_out, sl.n2o, self.max_nodes, sl.num_items, self.layout, self.stats,
                               self.remote)

    def gather_labels(self, seeds, n_seeds):
        """DoCPULabelExtractAndCopy, on the GPU (GPUExtract with D = 1, int64)."""
        K.row_copy(self.label_out, None, self.label_src, seeds, n_seeds, None, 8)

    def extract(self, seeds=None, n_seeds=0, slot=0):
        self.gather(slot)
        if self.label_src is not None and seeds is not None:
            self.gather_labels(seeds, n_seeds)

    def split(self, bufs, slot=0):
        """GetMissCacheIndex: explicit index lists (reference-compatible output)."""
        sl = self.slots[slot]
        K.cache_split(self.cache_table, sl.n2o, self.max_nodes, sl.num_items, bufs[0], bufs[1], bufs[2], bufs[3],
                      self.split_counts, sl.chain)

    def step(self, seeds, n_seeds, batch_key, slot=0):
        self.sample(seeds, n_seeds, batch_key, slot)
        self.extract(seeds, n_seeds, slot)
