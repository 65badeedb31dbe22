"""Feature cache partitioned across the trainer GPUs of one box (BASELINE north_star).

The reference replicates the cache in every trainer (dist_engine.cc:418-424) and has every trainer
re-gather all cached rows at init (dist_cache_manager_host.cc:98-109).  Here cache slot s (= position in the
PreSC ranking) lives on trainer `s % T` at local row `s // T`; each GPU fills only its own stripe, exports the
shard as a CUDA IPC handle, and maps the peers' shards, so fgnn_k_gather_cached reads remote rows with plain
loads over NVLink.  Two one-shot collectives at init, none on the data path:
  * the PreSC ranking is broadcast from rank 0 (replaces the shared-memory hop of dist_engine.cc:119-123),
  * the 64-byte IPC handles are all-gathered.
The host logic (striping, handle exchange, broadcast) is backend-agnostic and runs under gloo on CPU in
tests/test_dist_cpu.py; the device side needs the CUDA extension (no fallback).
"""
import torch
import torch.distributed as dist

from . import kernels as K


def stripe_rows(num_cached, num_shards, shard_id):
    """Number of cache slots owned by `shard_id`: slots shard_id, shard_id+T, ... < num_cached."""
    if num_cached <= shard_id:
        return 0
    return (num_cached - shard_id + num_shards - 1) // num_shards


def slot_owner(slot, num_shards, num_replicated=0):
    """(owner, local_row) of a cache slot; must match RowSrc::resolve in csrc/kernels/gather.cu.  Slots below
    `num_replicated` are held by every GPU (owner None); the slots behind them are striped."""
    if slot < num_replicated:
        return None, slot
    s = slot - num_replicated
    return s % num_shards, s // num_shards


def _comm_device(device):
    return torch.device(device) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_ranking(ranking_nodes, src=0):
    """One-shot broadcast of the PreSC ranking (u32[V] as int32 bits) from `src` to every rank, in place."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(ranking_nodes, src=src)
    return ranking_nodes


def allreduce_freq(freq):
    """Sum the per-rank PreSC visit counters (each rank pre-samples its own share of the epoch)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(freq, op=dist.ReduceOp.SUM)
    return freq


def exchange_handles(handle, device="cpu"):
    """All-gather one opaque 64-byte handle per rank; returns the list ordered by rank."""
    assert len(handle) == 64
    world = dist.get_world_size()
    dev = _comm_device(device)
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    out = [torch.empty(64, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


class CacheShards:
    """This rank's stripe of the feature cache plus the peer mappings of all the others."""

    def __init__(self, ranking_nodes, num_cached, feat_src, row_bytes, feat_mask, rank, world, device,
                 num_replicated=0):
        self.rank, self.world, self.row_bytes = rank, world, row_bytes
        # hybrid layout: the hottest `num_replicated` ranks on every GPU, the tail striped over the GPUs
        self.num_replicated = R = min(int(num_replicated), num_cached)
        self.replica_ptr, self.replica_bytes = None, 0
        if R:
            self.replica_bytes = R * row_bytes
            self.replica_ptr = K.shard_alloc(self.replica_bytes)
            K.row_copy(self.replica_ptr, None, feat_src, ranking_nodes[:R].contiguous(), R, None, row_bytes, feat_mask)
        self.local_rows = stripe_rows(num_cached - R, world, rank)
        nbytes = max(1, self.local_rows) * row_bytes
        self.ptr = K.shard_alloc(nbytes)                       # plain cudaMalloc: exportable
        if self.local_rows:
            idx = ranking_nodes[R + rank:num_cached:world].contiguous()
            K.row_copy(self.ptr, None, feat_src, idx, self.local_rows, None, row_bytes, feat_mask)
        torch.cuda.synchronize()
        handles = exchange_handles(K.ipc_export(self.ptr), device)
        self.opened = []
        ptrs = []
        for r in range(world):
            if r == rank:
                ptrs.append(self.ptr)
            else:
                p = K.ipc_open(handles[r])                     # NVLink peer mapping of rank r's shard
                self.opened.append(p)
                ptrs.append(p)
        self.ptrs = ptrs
        self.nbytes = nbytes

    def close(self):
        torch.cuda.synchronize()
        for p in self.opened:
            K.ipc_close(p)
        self.opened = []
        if dist.is_initialized():
            dist.barrier()                                     # nobody frees while a peer still maps it
        if self.ptr:
            K.shard_free(self.ptr)
            self.ptr = 0
        if self.replica_ptr:
            K.shard_free(self.replica_ptr)
            self.replica_ptr = None
