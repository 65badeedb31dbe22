"""arch5 sampler -> trainer queue: threaded stress test of the ticket protocol (csrc/runtime/rt_ring.h; reference
MemoryQueue, memory_queue.cc:51-138) on CPU.  Several producers, several consumers that hold their slots for random
times, so slots are released out of order — the situation of S samplers feeding T >= 2 trainers."""
import ctypes
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fgnn-artifacts_b200", "samgraph", "torch", "c_lib.so")


@pytest.fixture(scope="module")
def selftest():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(LIB)
    f = lib.fgnn_rt_ring_selftest
    f.argtypes = [ctypes.c_uint32] * 4 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
    f.restype = ctypes.c_long
    return f


CASES = [  # num_slots, slot_words, producers, consumers, items, max consumer delay (us)
    (3, 4096, 1, 2, 400, 300),     # the failing GPU scenario: 1 sampler, 2 trainers, max_copying_jobs + 1 = 3 slots
    (2, 1024, 2, 3, 600, 200),
    (4, 256, 3, 2, 2000, 50),
    (16, 64, 4, 4, 5000, 20),      # 16 slots = the engine's upper bound
    (3, 4096, 1, 1, 300, 100),     # single consumer: in-order releases
    (2, 8, 8, 8, 20000, 0),        # no delay: pure contention
    (5, 128, 2, 2, 0, 10),         # nothing to do
]


@pytest.mark.parametrize("case", CASES)
def test_every_record_arrives_once_and_intact(selftest, case):
    ns, words, P, C, items, delay = case
    assert selftest(ns, words, P, C, items, delay, 60000, 0) == 0


def test_negative_control_fill_level_alone_is_not_enough(selftest):
    """Without the per-slot wait (what the engine did before this test existed: only the fill-level semaphore, like
    a single-consumer ring) out-of-order releases corrupt records or dead-lock; the stress test must see it."""
    outcomes = [selftest(ns, words, P, C, items, delay, 3000, 1) for (ns, words, P, C, items, delay) in CASES[:3]]
    assert any(o != 0 for o in outcomes), outcomes
