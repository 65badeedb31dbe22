import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fgnn-artifacts_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout(seconds): per-test limit (pytest-timeout; a no-op marker without the plugin)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import RefCPU, build_ref, have_ref
    try:
        build_ref()
    except Exception:
        pass
    if not have_ref():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return RefCPU()


def small_graph(num_nodes=3000, num_edges=60000, seed=7, zero_deg_frac=0.05):
    """Power-law test graph with isolated vertices, low- and high-degree rows."""
    from fgnn_b200.synth import make_graph_numpy
    indptr, indices = make_graph_numpy(num_nodes, num_edges, seed)
    # carve out isolated vertices (the reference's len == 0 paths)
    rng = np.random.default_rng(seed)
    deg = np.diff(indptr.astype(np.int64))
    kill = rng.random(num_nodes) < zero_deg_frac
    keep_edge = np.repeat(~kill, deg)
    deg[kill] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    indices = np.ascontiguousarray(indices[keep_edge])
    return indptr, indices


@pytest.fixture(scope="session")
def graph_small():
    return small_graph()


@pytest.fixture(scope="session")
def graph_medium():
    return small_graph(1 << 16, 1 << 20, seed=11)
