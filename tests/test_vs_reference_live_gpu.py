"""BASELINE config #1 on the GPU, against the reference's OWN code run live (oracle/_ref, shipped prebuilt to the GPU
box): GCN fanout [5,10,15], batch 8000, products-shaped graph.  The reference's CPU sampler draws the edges
(thread-local mt19937, cannot be replayed); the CUDA unique / remap / extraction path is fed exactly those edges and
must reproduce CPUHashTable2's unique list, MapEdges' local ids and CPUExtract's rows bit for bit (parity definition
(2) of SURVEY §8c) — no oracle in between."""
import numpy as np
import pytest

from test_kernels_gpu import HT, K, dev, host  # noqa: F401  (fixtures + helpers)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]   # pytest-timeout: a hang must not eat the GPU run
torch = pytest.importorskip("torch")


def test_gcn_products_reference_edges_through_cuda(K, ref):
    from fgnn_b200.synth import SHAPES, make_graph_numpy
    V, E, D, C, T = SHAPES["products"]
    scale = 32
    V, E = V // scale, E // scale
    fanout = [5, 10, 15]
    indptr, indices = make_graph_numpy(V, E, seed=31)
    rng = np.random.default_rng(4)
    feat = (rng.random((V, D), dtype=np.float32) * 2 - 1).astype(np.float32)
    label = rng.integers(0, C, size=V).astype(np.int64)
    seeds = rng.permutation(V)[:8000].astype(np.uint32)
    ref.set_threads(1)

    rh = ref.hashtable(2, V)
    ht = HT(K, V + 16)
    rh.populate(seeds)
    ht.fill_unique(seeds)
    cur = seeds
    work = indices.copy()
    for i in (2, 1, 0):
        src, dst = ref.sample_khop2(indptr, work, cur, fanout[i])          # the reference's sampled edges
        n = len(dst)
        rh.populate(dst)
        d_dst, pos = ht.fill_duplicates(dst)
        unique_ref = rh.map_nodes()
        assert ht.num_items() == len(unique_ref)
        assert np.array_equal(ht.unique(), unique_ref), "layer %d: ordered unique list differs from CPUHashTable2" % i
        new_src_ref, new_dst_ref = rh.map_edges(src, dst)
        assert np.array_equal(ht.map(None, pos, n), new_dst_ref)            # neighbours, by remembered bucket
        assert np.array_equal(ht.map(dev(src), None, n), new_src_ref)       # seeds, by probing
        cur = unique_ref

    # extraction of the batch's input nodes and labels: GPUExtract == CPUExtract
    n_in = len(cur)
    d_feat = torch.from_numpy(feat).cuda()
    out = torch.zeros((n_in, D), dtype=torch.float32, device="cuda")
    K.row_copy(out, None, d_feat, dev(cur), n_in, None, D * 4)
    d_label = torch.from_numpy(label).cuda()
    lab = torch.zeros(len(seeds), dtype=torch.int64, device="cuda")
    K.row_copy(lab, None, d_label, dev(seeds), len(seeds), None, 8)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref.extract(feat, cur).view(np.uint32))
    assert np.array_equal(lab.cpu().numpy(), ref.extract(label, seeds))

    # the same rows through the cache-aware gather (30 % of the vertices cached in HBM, the rest read from pinned
    # host memory): one kernel instead of split + CPU gather + H2D + two combines, same bytes
    nc = int(V * 0.3)
    rank = torch.from_numpy(rng.permutation(V).astype(np.int32)).cuda()
    ctable = torch.empty(V, dtype=torch.int32, device="cuda")
    K.cache_table_build(ctable, V, rank, nc)
    cache = torch.empty((nc, D), dtype=torch.float32, device="cuda")
    K.row_copy(cache, None, d_feat, rank, nc, None, D * 4)
    host_feat = torch.from_numpy(feat).pin_memory()
    out2 = torch.zeros((n_in + 7, D), dtype=torch.float32, device="cuda")
    ptrs = torch.tensor([cache.data_ptr()], dtype=torch.int64, device="cuda")
    d_n = torch.tensor([n_in], dtype=torch.int32, device="cuda")
    K.gather_cached(out2, dev(cur), n_in, d_n, ctable, ptrs, 1, host_feat, D * 4)
    torch.cuda.synchronize()
    assert np.array_equal(out2[:n_in].cpu().numpy().view(np.uint32), ref.extract(feat, cur).view(np.uint32))
    assert float(out2[n_in:].abs().sum().item()) == 0.0
