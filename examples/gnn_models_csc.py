"""DGL-free GCN and PinSAGE consumers of the CSC block hand-off (sam.get_csc_blocks), next to the GraphSAGE model of
train_graphsage_csc.py: the three model families the reference trains (example/samgraph/multi_gpu/train_gcn.py,
train_graphsage.py, train_pinsage.py).  Aggregation is one sparse-CSR SpMM per layer.

  GCN       dgl.nn.GraphConv(norm='both', allow_zero_in_degree=True) as used by train_gcn.py:18-47:
            h_dst' = D_in^-1/2 · A · D_out^-1/2 · h_src · W + b, degrees counted inside the block and clamped to >= 1
  PinSAGE   WeightedSAGEConv of train_pinsage.py:30-66: n = relu(Q·dropout(h_src)); per destination the sum of
            w_e · n_src over its edges divided by max(sum of w_e, 1) (w = the visit counts the random-walk sampler
            emits as edge data); z = relu(W·dropout([n/ws, h_dst])), rows scaled to unit L2 norm

Blocks are (indptr i32[num_dst+1], indices i32[E], num_src, num_dst) for GCN and the same plus `weights`[E] (in CSC
edge order) for PinSAGE.  Unit-tested on CPU against index_add formulations (tests/test_example_model_cpu.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _csr(indptr, indices, vals, num_dst, num_src):
    return torch.sparse_csr_tensor(indptr, indices, vals, size=(num_dst, num_src))


class GraphConvCSC(nn.Module):
    def __init__(self, in_feats, out_feats, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = nn.Parameter(torch.zeros(out_feats))
        nn.init.xavier_uniform_(self.weight)
        self.activation = activation
        self.in_feats, self.out_feats = in_feats, out_feats

    def forward(self, block, h):
        indptr, indices, num_src, num_dst = block
        assert h.shape[0] == num_src
        counts = (indptr[1:] - indptr[:-1]).long()
        in_deg = counts.to(h.dtype).clamp(min=1)
        out_deg = torch.bincount(indices.long(), minlength=num_src).to(h.dtype).clamp(min=1)
        h = h * out_deg.pow(-0.5)[:, None]
        if self.in_feats > self.out_feats:          # GraphConv multiplies first when that shrinks the rows
            h = h @ self.weight
        adj = _csr(indptr, indices, torch.ones(indices.shape[0], dtype=h.dtype, device=h.device), num_dst, num_src)
        rst = torch.sparse.mm(adj, h)
        if self.in_feats <= self.out_feats:
            rst = rst @ self.weight
        rst = rst * in_deg.pow(-0.5)[:, None] + self.bias
        return self.activation(rst) if self.activation is not None else rst


class GCN(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, dropout, activation=F.relu):
        super().__init__()
        dims = [in_feats] + [n_hidden] * (n_layers - 1) + [n_classes]
        self.layers = nn.ModuleList(
            GraphConvCSC(dims[i], dims[i + 1], activation if i != n_layers - 1 else None) for i in range(n_layers))
        self.dropout = nn.Dropout(dropout)

    def forward(self, blocks, x):
        h = x
        for i, (layer, block) in enumerate(zip(self.layers, blocks)):
            if i != 0:
                h = self.dropout(h)                 # train_gcn.py:41-46
            h = layer(block[:4], h)
        return h


class WeightedSAGEConvCSC(nn.Module):
    def __init__(self, input_dims, hidden_dims, output_dims, dropout, act=F.relu):
        super().__init__()
        self.act = act
        self.Q = nn.Linear(input_dims, hidden_dims)
        self.W = nn.Linear(input_dims + hidden_dims, output_dims)
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.Q.weight, gain=gain)
        nn.init.xavier_uniform_(self.W.weight, gain=gain)
        nn.init.constant_(self.Q.bias, 0)
        nn.init.constant_(self.W.bias, 0)
        self.dropout = nn.Dropout(dropout)

    def forward(self, block, h):
        indptr, indices, num_src, num_dst, weights = block
        assert h.shape[0] == num_src
        h_dst = h[:num_dst]
        n = self.act(self.Q(self.dropout(h)))
        w = weights.to(n.dtype)
        agg = torch.sparse.mm(_csr(indptr, indices, w, num_dst, num_src), n)
        counts = (indptr[1:] - indptr[:-1]).long()
        dst_of = torch.repeat_interleave(torch.arange(num_dst, device=h.device), counts)
        ws = torch.zeros(num_dst, dtype=n.dtype, device=h.device).index_add_(0, dst_of, w).clamp(min=1)
        z = self.act(self.W(self.dropout(torch.cat([agg / ws[:, None], h_dst], 1))))
        z_norm = z.norm(2, 1, keepdim=True)
        z_norm = torch.where(z_norm == 0, torch.ones_like(z_norm), z_norm)
        return z / z_norm


class PinSAGE(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, dropout, activation=F.relu):
        super().__init__()
        dims = [in_feats] + [n_hidden] * (n_layers - 1) + [n_classes]
        self.layers = nn.ModuleList(
            WeightedSAGEConvCSC(dims[i], n_hidden, dims[i + 1], dropout, activation) for i in range(n_layers))

    def forward(self, blocks, x):
        h = x
        for layer, block in zip(self.layers, blocks):
            h = layer(block, h)
        return h


def csc_blocks_weighted(sam, batch_key, num_layers):
    """csc_blocks of train_graphsage_csc.py plus each layer's edge data (the random-walk visit counts,
    adapter.py:104-118 `edata['weights']`) in CSC edge order."""
    blocks, feat, label = sam.get_csc_blocks(batch_key, num_layers)
    out = []
    for i, (indptr, indices, eids, num_src, num_dst) in enumerate(blocks):
        w = sam.get_graph_data(batch_key, i)
        if eids is not None:
            w = w[eids.long()]
        out.append((indptr, indices, num_src, num_dst, w))
    return out, feat, label


def build_model(name, in_feats, n_hidden, n_classes, n_layers, dropout):
    if name == "gcn":
        return GCN(in_feats, n_hidden, n_classes, n_layers, dropout)
    if name == "pinsage":
        return PinSAGE(in_feats, n_hidden, n_classes, n_layers, dropout)
    raise ValueError(name)
