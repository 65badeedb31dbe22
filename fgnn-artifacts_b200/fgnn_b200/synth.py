"""Deterministic synthetic graphs of the reference's dataset shapes (no datasets
are available offline).  Shapes come from the reference's meta.txt writers:
datagen/products.py:88-94, datagen/papers100M.py:86-95, datagen/twitter.sh:35-43,
datagen/uk-2006-05.sh:35-43.

Law (with one change): in-degree of vertex pi(r) ~ r^-s over ranks r,
scaled so the degrees sum to E and clamped to 2^20; neighbour ids drawn i.i.d.
from Zipf(s=0.9) over a second permutation (popular sources -> non-trivial cache
hit rates).  CSR of in-neighbours, uint32 ids (stored in int32 tensors on torch).
The degree exponent is s=0.5 instead of the survey's 1.0: with s=1.0 the median
degree of a papers100M-shaped graph is 2 and a GraphSAGE [25,10] batch samples
only ~0.15 M edges, an order of magnitude less work than the ~1-1.5 M edges/batch
the reference reports on the real dataset; s=0.5 (median degree
~10, max ~77 k) lands at ~1 M edges and ~0.7 M input nodes per batch.

`make_graph_numpy` is used by the CPU tests, `make_graph_torch` generates the
same law on the GPU for the full-size benchmark shapes.
"""
import math

import numpy as np

SEED = 0x46474E4E  # "FGNN"

SHAPES = {
    # name: (num_node, num_edge, feat_dim, num_class, num_train)
    "products": (2449029, 123718152, 100, 47, 196615),
    "papers100M": (111059956, 1615685872, 128, 172, 1207179),
    "twitter": (41652230, 1468365182, 256, 150, 416500),
    "uk-2006-05": (77741046, 2965197340, 256, 150, 1000000),
    # small CI shapes with the same law
    "ci-64k": (1 << 16, 1 << 20, 32, 16, 4096),
    "ci-1m": (1 << 20, 1 << 24, 128, 32, 65536),
}

MAX_DEG = 1 << 20


def _degree_scale(num_nodes, num_edges, s=1.0):
    """c such that sum_r min(c * r^-s, MAX_DEG) == num_edges (bisection)."""
    r = np.arange(1, num_nodes + 1, dtype=np.float64)
    w = r ** (-s)
    lo, hi = 0.0, float(num_edges) * 4
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if np.minimum(mid * w, MAX_DEG).sum() < num_edges:
            lo = mid
        else:
            hi = mid
    return hi


def make_graph_numpy(num_nodes, num_edges, seed=SEED, s_deg=0.5, s_nbr=0.9):
    rng = np.random.Generator(np.random.Philox(seed))
    c = _degree_scale(num_nodes, num_edges, s_deg)
    r = np.arange(1, num_nodes + 1, dtype=np.float64)
    degf = np.minimum(c * r ** (-s_deg), MAX_DEG)
    perm = rng.permutation(num_nodes)
    degv = np.empty(num_nodes, np.float64)
    degv[perm] = degf
    cum = np.concatenate([[0.0], np.cumsum(degv)])
    indptr = np.rint(cum * (num_edges / cum[-1])).astype(np.uint64)
    indptr[-1] = num_edges
    indptr = indptr.astype(np.uint32)
    perm2 = rng.permutation(num_nodes).astype(np.uint32)
    u = rng.random(num_edges)
    a = 1.0 - s_nbr
    rank = np.floor((1.0 + u * (num_nodes ** a - 1.0)) ** (1.0 / a)).astype(np.int64) - 1
    rank = np.clip(rank, 0, num_nodes - 1)
    indices = perm2[rank]
    return indptr, indices


def make_dataset_numpy(name_or_shape, seed=SEED, with_feat=True):
    shape = SHAPES[name_or_shape] if isinstance(name_or_shape, str) else name_or_shape
    V, E, D, C, T = shape
    indptr, indices = make_graph_numpy(V, E, seed)
    rng = np.random.Generator(np.random.Philox(seed + 1))
    ds = dict(num_node=V, num_edge=E, feat_dim=D, num_class=C, indptr=indptr, indices=indices)
    if with_feat:
        ds["feat"] = (rng.random((V, D), dtype=np.float32) * 2.0 - 1.0).astype(np.float32)
    ds["label"] = rng.integers(0, C, size=V, dtype=np.int64)
    ds["train_set"] = rng.permutation(V)[:T].astype(np.uint32)
    wrng = np.random.Generator(np.random.Philox(seed + 2))
    ds["edge_weight"] = wrng.integers(1, 11, size=E).astype(np.float32)  # create_alias_table.cc:113
    return ds


def make_graph_torch(num_nodes, num_edges, device="cuda", seed=SEED, s_deg=0.5, s_nbr=0.9,
                     chunk=1 << 28):
    """Same law, generated on `device`.  Returns int32-typed (uint32 bits) indptr, indices."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    c = _degree_scale_fast(num_nodes, num_edges, s_deg)
    r = torch.arange(1, num_nodes + 1, dtype=torch.float64, device=device)
    degf = torch.clamp(c * r.pow(-s_deg), max=float(MAX_DEG))
    del r
    perm = torch.randperm(num_nodes, generator=g, device=device)
    degv = torch.empty_like(degf)
    degv[perm] = degf
    del perm, degf
    cum = torch.cumsum(degv, 0)
    del degv
    total = float(cum[-1].item())
    indptr64 = torch.zeros(num_nodes + 1, dtype=torch.int64, device=device)
    indptr64[1:] = torch.round(cum * (num_edges / total)).to(torch.int64)
    indptr64[-1] = num_edges
    del cum
    indptr = _to_u32_bits(indptr64)
    del indptr64
    perm2 = torch.randperm(num_nodes, generator=g, device=device).to(torch.int32)
    indices = torch.empty(num_edges, dtype=torch.int32, device=device)
    a = 1.0 - s_nbr
    top = num_nodes ** a - 1.0
    for lo in range(0, num_edges, chunk):
        hi = min(num_edges, lo + chunk)
        u = torch.rand(hi - lo, generator=g, device=device, dtype=torch.float64)
        rank = torch.floor((1.0 + u * top).pow(1.0 / a)).to(torch.int64) - 1
        rank.clamp_(0, num_nodes - 1)
        indices[lo:hi] = perm2[rank]
        del u, rank
    return indptr, indices


def _degree_scale_fast(num_nodes, num_edges, s=1.0):
    """Bisection on the closed-form partial sums (no V-sized temporaries)."""
    def total(c):
        # ranks r <= r_cap are clamped: c * r^-s >= MAX_DEG  <=>  r <= (c/MAX_DEG)^(1/s)
        r_cap = min(num_nodes, int((c / MAX_DEG) ** (1.0 / s))) if c > MAX_DEG else 0
        head = r_cap * float(MAX_DEG)
        # sum_{r=r_cap+1..V} r^-s  ~  integral + Euler-Maclaurin first term
        a, b = r_cap + 1, num_nodes
        if a > b:
            return head
        if abs(s - 1.0) < 1e-12:
            tail = math.log(b / a) + 0.5 * (1.0 / a + 1.0 / b)
        else:
            tail = (b ** (1 - s) - a ** (1 - s)) / (1 - s) + 0.5 * (a ** -s + b ** -s)
        return head + c * tail
    lo, hi = 0.0, float(num_edges) * 4
    for _ in range(100):
        mid = 0.5 * (lo + hi)
        if total(mid) < num_edges:
            lo = mid
        else:
            hi = mid
    return hi


def _to_u32_bits(t64):
    """int64 values in [0, 2^32) -> int32 tensor holding the same low 32 bits."""
    import torch

    return torch.where(t64 >= (1 << 31), t64 - (1 << 32), t64).to(torch.int32)


def u32_tensor(np_u32, device="cuda"):
    """numpy uint32 -> torch int32 (same bits) on device."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(np_u32).view(np.int32)).to(device)


def to_np_u32(t):
    return t.detach().cpu().numpy().view(np.uint32)


def write_dataset(path, ds, with_weights=False, oracle=None):
    """Write `ds` (dict from make_dataset_numpy) in the reference's on-disk format
    (datagen/README.md, engine.cc:73-264, constant.cc:23-50): meta.txt + *.bin."""
    import os

    os.makedirs(path, exist_ok=True)
    V, E = ds["num_node"], ds["num_edge"]
    n_test = n_valid = min(1000, V // 10)
    rng = np.random.default_rng(5)
    test = rng.permutation(V)[:n_test].astype(np.uint32)
    valid = rng.permutation(V)[:n_valid].astype(np.uint32)
    with open(os.path.join(path, "meta.txt"), "w") as f:
        f.write("NUM_NODE %d\nNUM_EDGE %d\nFEAT_DIM %d\nNUM_CLASS %d\nNUM_TRAIN_SET %d\nNUM_VALID_SET %d\nNUM_TEST_SET %d\n"
                % (V, E, ds["feat_dim"], ds["num_class"], len(ds["train_set"]), n_valid, n_test))
    ds["indptr"].astype(np.uint32).tofile(os.path.join(path, "indptr.bin"))
    ds["indices"].astype(np.uint32).tofile(os.path.join(path, "indices.bin"))
    if "feat" in ds:
        ds["feat"].astype(np.float32).tofile(os.path.join(path, "feat.bin"))
    ds["label"].astype(np.uint64).tofile(os.path.join(path, "label.bin"))
    ds["train_set"].astype(np.uint32).tofile(os.path.join(path, "train_set.bin"))
    test.tofile(os.path.join(path, "test_set.bin"))
    valid.tofile(os.path.join(path, "valid_set.bin"))
    if with_weights:
        assert oracle is not None, "weight tables are built by the oracle's restatement of the reference tools"
        prob, alias = oracle.build_alias_table(ds["indptr"], ds["indices"], ds["edge_weight"])
        prefix = oracle.build_prefix_table(ds["indptr"], ds["edge_weight"])
        prob.tofile(os.path.join(path, "prob_table.bin"))
        alias.tofile(os.path.join(path, "alias_table.bin"))
        prefix.tofile(os.path.join(path, "prob_prefix_table.bin"))
        ds["prob_table"], ds["alias_table"], ds["prob_prefix_table"] = prob, alias, prefix
    return path
