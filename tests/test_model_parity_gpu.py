"""Model-output parity (BASELINE north_star: "model outputs must match within fp32 tolerance, 1e-5 relative").

A 2-layer mean-aggregator GraphSAGE forward (the model of example/samgraph/train_graphsage.py, written with plain
index_add so that neither DGL nor PyG is needed) is evaluated
  (a) on the GPU, in fp32, on the blocks + features the CUDA hot path produced (HotPath: fgnn_k_sample_batch +
      fgnn_k_gather_cached with a 30 % PreSC-style cache), once from the COO blocks and once from the CSC hand-off;
  (b) on the CPU, in fp64, on the blocks + features of the oracle for the same seeds and RNG counters.
Blocks and features are bit-identical (checked), so the only difference left is fp32 vs fp64 arithmetic:
max |a - b| <= 1e-5 * max |b|, the tolerance BASELINE.json states."""
import numpy as np
import pytest

from test_kernels_gpu import K, G, dev, gm, host, pick_seeds  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SEED, KEY = 0xC0FFEE, 9
FANOUTS = [10, 5]           # samgraph order: graphs[0] is the input-side (largest) block
D, HIDDEN, CLASSES = 32, 64, 16
TOL = 1e-5


def sage_forward(blocks, feat, weights, xp):
    """blocks[i] = (row=src local, col=dst local, num_src, num_dst); dst nodes are a prefix of src nodes
    (cuda_loops.cc:214-221).  h_dst' = act(W_self h_dst + W_neigh mean_{src->dst} h_src + b)."""
    h = feat
    for i, (row, col, num_src, num_dst) in enumerate(blocks):
        w_self, w_neigh, b = weights[i]
        assert h.shape[0] == num_src
        agg, deg = xp.scatter_sum(h, row, col, num_dst)
        mean = agg / xp.clamp_min1(deg)[:, None]
        h = h[:num_dst] @ w_self + mean @ w_neigh + b
        if i + 1 < len(blocks):
            h = xp.relu(h)
    return h


class TorchOps:
    @staticmethod
    def scatter_sum(h, row, col, num_dst):
        agg = torch.zeros((num_dst, h.shape[1]), dtype=h.dtype, device=h.device)
        agg.index_add_(0, col, h[row])
        deg = torch.zeros(num_dst, dtype=h.dtype, device=h.device)
        deg.index_add_(0, col, torch.ones(col.shape[0], dtype=h.dtype, device=h.device))
        return agg, deg

    clamp_min1 = staticmethod(lambda d: d.clamp(min=1))
    relu = staticmethod(torch.relu)


class NumpyOps:
    @staticmethod
    def scatter_sum(h, row, col, num_dst):
        agg = np.zeros((num_dst, h.shape[1]), h.dtype)
        np.add.at(agg, col, h[row])
        deg = np.bincount(col, minlength=num_dst).astype(h.dtype)
        return agg, deg

    clamp_min1 = staticmethod(lambda d: np.maximum(d, 1))
    relu = staticmethod(lambda x: np.maximum(x, 0))


def test_graphsage_outputs_match_oracle_pipeline(K, oracle, gm):
    from fgnn_b200.pipeline import HotPath
    from oracle.oracle import sample_batch_oracle
    V = len(gm.indptr_np) - 1
    rng = np.random.default_rng(3)
    feat = (rng.random((V, D), dtype=np.float32) * 2 - 1).astype(np.float32)
    seeds = pick_seeds(gm.indptr_np, 1500, 21)
    L = len(FANOUTS)

    # ---- oracle pipeline (CPU) ----
    exp = sample_batch_oracle(oracle, dict(indptr=gm.indptr_np, indices=gm.indices_np), seeds, FANOUTS, "khop2", SEED, KEY)
    feat_cpu = oracle.extract(feat, exp["input_nodes"])

    # ---- CUDA hot path ----
    hp = HotPath(gm.indptr, gm.indices, V, FANOUTS, len(seeds), "khop2", seed=SEED)
    d_seeds = dev(seeds)
    hp.sample(d_seeds, len(seeds), KEY)
    freq = torch.zeros(V, dtype=torch.int32, device="cuda")
    hp.presample_count(freq)
    rank = torch.empty(V, dtype=torch.int32, device="cuda")
    ws = torch.empty(K.presc_rank_workspace_bytes(V), dtype=torch.uint8, device="cuda")
    K.presc_rank(freq, V, rank, ws)
    host_feat = torch.from_numpy(feat).pin_memory()
    hp.build_cache(rank, 0.3, host_feat, D * 4)
    hp.gather()
    torch.cuda.synchronize()
    counts = hp.counts.cpu().numpy()
    n_in = int(hp.num_items.item())
    assert np.array_equal(host(hp.n2o, n_in), exp["input_nodes"])
    feat_gpu = hp.feat_out[:n_in].view(torch.float32).reshape(n_in, D)
    assert np.array_equal(feat_gpu.cpu().numpy().view(np.uint32), feat_cpu.view(np.uint32))   # extraction bit-exact

    blocks_gpu, blocks_csc, blocks_cpu = [], [], []
    for i in range(L):
        num_dst, num_edge, num_src = (int(x) for x in counts[i])
        e = exp["layers"][i]
        assert (num_dst, num_edge, num_src) == (e["num_dst"], e["num_edge"], e["num_src"])
        row, col = hp.row[i][:num_edge], hp.col[i][:num_edge]
        assert np.array_equal(host(row), e["row"]) and np.array_equal(host(col), e["col"])    # blocks bit-exact
        blocks_gpu.append((row.long(), col.long(), num_src, num_dst))
        # the same block through the CSC hand-off (indptr + zero-copy indices): expand indptr back to dst ids
        indptr = torch.empty(num_dst + 1, dtype=torch.int32, device="cuda")
        K.coo_to_csc(row, col, num_edge, None, num_dst, True, indptr)
        dst_of = torch.repeat_interleave(torch.arange(num_dst, device="cuda"), torch.diff(indptr.long()))
        blocks_csc.append((row.long(), dst_of, num_src, num_dst))
        blocks_cpu.append((e["row"].astype(np.int64), e["col"].astype(np.int64), num_src, num_dst))

    dims = [D, HIDDEN, CLASSES]
    wrng = np.random.default_rng(17)
    weights64 = [(wrng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i]),
                  wrng.standard_normal((dims[i], dims[i + 1])) / np.sqrt(dims[i]),
                  wrng.standard_normal(dims[i + 1]) * 0.1) for i in range(L)]
    weights_gpu = [tuple(torch.from_numpy(w.astype(np.float32)).cuda() for w in ws_) for ws_ in weights64]

    ref64 = sage_forward(blocks_cpu, feat_cpu.astype(np.float64), weights64, NumpyOps)
    assert ref64.shape == (len(seeds), CLASSES)
    scale = float(np.abs(ref64).max())
    for name, blocks in (("coo", blocks_gpu), ("csc", blocks_csc)):
        out = sage_forward(blocks, feat_gpu, weights_gpu, TorchOps)
        torch.cuda.synchronize()
        err = float(np.abs(out.double().cpu().numpy() - ref64).max())
        assert err <= TOL * scale, "%s blocks: max abs err %.3e vs tolerance %.3e" % (name, err, TOL * scale)
