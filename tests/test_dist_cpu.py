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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    steps_per_epoch, K = 151, 40
    mine = [bench.step_of(k, rank, world, steps_per_epoch) for k in range(K)]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ms, edges = bench.aggregate(10.0 + rank, 1000 * (rank + 1), device="cpu")
    ret[rank] = (gathered, ms, edges)
    dist.destroy_process_group()


def test_rank_sharding_and_aggregation_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for rank in range(world):
        gathered, ms, edges = ret[rank]
        flat = [s for g in gathered for s in g]
        assert len(set(flat)) == len(flat), "ranks must work on disjoint mini-batches"
        assert ms == 11.0 and edges == 3000          # max over ranks, sum over ranks
