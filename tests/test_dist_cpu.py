"""World-size-2 gloo test (CPU) of the N>1 host logic: the split of an epoch's mini-batches over ranks
(no data-path collective) and the whole-job aggregation used by bench.py (max time, summed work)."""
import os
import socket
import sys

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    steps_per_epoch, K = 151, 40
    mine = [bench.step_of(k, rank, world, steps_per_epoch) for k in range(K)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms, edges = bench.aggregate(10.0 + rank, 1000 * (rank + 1), device="cpu")
    # partitioned-cache init plumbing (fgnn_b200/partition.py): PreSC counters summed, ranking broadcast from
    # rank 0, one opaque 64-byte IPC handle per rank all-gathered in rank order
    from fgnn_b200 import partition as P
    freq = torch.arange(16, dtype=torch.int32) * (rank + 1)
    P.allreduce_freq(freq)
    ranking = torch.arange(16, dtype=torch.int32).flip(0) if rank == 0 else torch.zeros(16, dtype=torch.int32)
    P.broadcast_ranking(ranking, src=0)
    handles = P.exchange_handles(bytes([rank + 1] * 64), "cpu")
    ret[rank] = (gathered, ms, edges, freq.tolist(), ranking.tolist(), handles)
    dist.destroy_process_group()


def test_rank_sharding_and_aggregation_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        gathered, ms, edges, freq, ranking, handles = ret[rank]
        assert freq == [3 * i for i in range(16)]
        assert ranking == list(range(15, -1, -1))
        assert handles == [bytes([r + 1] * 64) for r in range(world)]
        flat = [s for g in gathered for s in g]
        assert len(set(flat)) == len(flat), "ranks must work on disjoint mini-batches"
        assert ms == 11.0 and edges == 3000          # max over ranks, sum over ranks


def test_cache_striping_covers_every_slot_once():
    sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))
    from fgnn_b200 import partition as P
    for num_cached in (0, 1, 7, 64, 1001):
        for T in (1, 2, 3, 8):
            rows = [P.stripe_rows(num_cached, T, t) for t in range(T)]
            assert sum(rows) == num_cached
            seen = set()
            for s in range(num_cached):
                o, r = P.slot_owner(s, T)
                assert 0 <= o < T and r < rows[o]
                seen.add((o, r))
            assert len(seen) == num_cached


def test_hybrid_layout_replicated_head_and_striped_tail():
    """fgnn_cache_layout on the host side: slots below the replicated head belong to every GPU (owner None, row =
    slot), the slots behind it are striped: every one of them has exactly one owner and a unique local row, and the
    stripes' sizes are those the trainers allocate (Extractor ctor / partition.CacheShards)."""
    sys.path.insert(0, os.path.join(ROOT, "fgnn-artifacts_b200"))
    from fgnn_b200 import partition as P
    for num_cached in (0, 5, 64, 1001):
        for T in (1, 2, 3, 8):
            for R in (0, 1, num_cached // 4, num_cached, num_cached + 7):
                head = min(R, num_cached)
                rows = [P.stripe_rows(num_cached - head, T, t) for t in range(T)]
                assert sum(rows) == num_cached - head
                seen = set()
                for s in range(num_cached):
                    o, r = P.slot_owner(s, T, head)
                    if s < head:
                        assert o is None and r == s
                    else:
                        assert 0 <= o < T and r < rows[o]
                        seen.add((o, r))
                assert len(seen) == num_cached - head


def test_factored_split_follows_the_reference_placements():
    """bench.py's S+T split of the factored legs (multi_gpu/common_config.py:182-185; 2 samplers + 6 trainers on 8)."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.factored_split(1) == (1, 1, True)
    assert bench.factored_split(2) == (1, 1, False)
    assert bench.factored_split(4) == (1, 3, False)
    assert bench.factored_split(8) == (2, 6, False)
    for n in (2, 3, 4, 5, 6, 7, 8):
        s, t, single = bench.factored_split(n)
        assert s >= 1 and t >= 1 and s + t == n and not single
