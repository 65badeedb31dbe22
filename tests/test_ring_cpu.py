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


DEFERRED = [  # the arch5 shapes of the multi-GPU runs: slots = max_copying_jobs + 1 = 5
    (5, 512, 2, 6, 3000, 100),     # 2 samplers + 6 trainers on 8 GPUs (hung in round 2 with release-after-take)
    (5, 512, 1, 3, 2000, 100),     # 1 + 3 on 4 GPUs
    (3, 512, 2, 6, 2000, 50),
    (2, 64, 4, 8, 4000, 20),
]


@pytest.mark.parametrize("case", DEFERRED)
def test_deferred_slot_release_in_the_engines_order(selftest, case):
    """Engine::RecvTask releases a slot only when the copies out of it have completed, i.e. later than it takes the
    record; it does so BEFORE requesting the next ticket (mode 3)."""
    ns, words, P, C, items, delay = case
    assert selftest(ns, words, P, C, items, delay, 60000, 3) == 0


def test_negative_control_release_after_next_ticket_hangs(selftest):
    """Mode 2 = the order the engine had when 2 samplers + 6 trainers hung on 8 GPUs: a trainer that still holds
    slot (t mod N) takes ticket t + N and waits for its publication, which waits for that slot."""
    outcomes = []
    for attempt in range(8):                  # a race: each attempt hangs with high, not full, probability
        outcomes.append(selftest(5, 512, 2, 6, 20000, 100, 5000, 2))
        if outcomes[-1] == -1:
            break
    assert outcomes[-1] == -1, outcomes
    assert selftest(5, 512, 2, 6, 20000, 100, 60000, 3) == 0      # same load, the engine's order


def test_sanity_check_batch_flags_what_the_reference_asserts():
    """SAMGRAPH_SANITY_CHECK (cuda_shuffler.cc:144-151, cuda_sanity_check.cu:29-59): empty keys, out-of-range ids and
    a train node handed out twice within one epoch ("duplicate batch input")."""
    import numpy as np
    lib = ctypes.CDLL(LIB)
    f = lib.fgnn_rt_sanity_check_batch
    f.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    f.restype = ctypes.c_int
    V = 1000
    emap = np.zeros(V, np.uint8)
    bad = ctypes.c_size_t(0)

    def check(ids):
        a = np.ascontiguousarray(ids, np.uint32)
        return f(emap.ctypes.data, V, a.ctypes.data, len(a), ctypes.byref(bad))

    perm = np.random.default_rng(1).permutation(V).astype(np.uint32)
    assert check(perm[:400]) == 0 and check(perm[400:800]) == 0          # two batches of one epoch
    assert int(emap.sum()) == 800
    assert check(perm[790:810]) == 3 and bad.value == 0                   # already handed out this epoch
    assert check([perm[900], perm[901], perm[900]]) == 3 and bad.value == 2
    assert check([0xFFFFFFFF]) == 1 and check([V]) == 2
    emap[:] = 0                                                           # next epoch
    assert check(perm) == 0 and check([]) == 0


def test_ring_protocol_across_processes():
    """The same protocol with fork()ed producers / consumers over one MAP_SHARED region (process-shared mutex and
    semaphores, sequence words in shared memory) — the arch5 situation.  Run from a fresh single-threaded interpreter
    (fork in a multi-threaded pytest process would be unsafe)."""
    import subprocess
    import sys
    code = (
        "import ctypes\n"
        "lib = ctypes.CDLL(%r)\n"
        "f = lib.fgnn_rt_ring_selftest_procs\n"
        "f.argtypes = [ctypes.c_uint32] * 4 + [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32]\n"
        "f.restype = ctypes.c_long\n"
        "cases = [(3, 4096, 1, 2, 400, 300), (3, 1024, 2, 2, 800, 200), (4, 256, 3, 2, 2000, 50),\n"
        "         (16, 64, 4, 4, 5000, 20), (2, 8, 4, 4, 20000, 0), (5, 16, 2, 2, 0, 5)]\n"
        "print([int(f(*c, 60000)) for c in cases])\n" % LIB)
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == "[0, 0, 0, 0, 0, 0]", r.stdout
