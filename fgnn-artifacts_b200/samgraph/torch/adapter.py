"""samgraph.torch — the API the training scripts use (`import samgraph.torch as sam`).

Same functions as the reference's samgraph/torch/adapter.py:30-179.  The C++ engine hands out torch.Tensor
views of its own memory (csrc/runtime/pymodule.cc: DLPack capsule adopted zero-copy by torch.utils.dlpack inside
the extension), exactly what the reference's c_lib returns, so the reference's adapter.py binds to this c_lib
unmodified too; DGL is imported lazily so that sampling/extraction works on boxes without it."""
import torch


def _from_dlpack(t):
    return t           # c_lib returns tensors (round 1 returned capsules)

from samgraph.common import *  # noqa: F401,F403  (enum constants, sample_types, builtin_archs, ...)
from samgraph.common import SamGraphBasics
from samgraph.torch import c_lib

_basics = SamGraphBasics(__file__, "c_lib")

# every samgraph_* entry point under its short name (config, init, start, sample_once, ...)
for _name in ("config init start shutdown num_class feat_dim num_epoch steps_per_epoch get_next_batch "
              "get_graph_num_src get_graph_num_dst get_graph_num_edge sample_once log_step log_step_add "
              "log_epoch_add get_log_init_value get_log_step_value get_log_epoch_value report_init report_step "
              "report_step_average report_epoch report_epoch_average report_node_access trace_step_begin "
              "trace_step_end trace_step_begin_now trace_step_end_now dump_trace forward_barrier wait_one_child "
              "switch_init data_init sample_init train_init extract_start num_local_step").split():
    globals()[_name] = getattr(_basics, _name)


def get_graph_feat(batch_key):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_feat(batch_key))


def get_graph_label(batch_key):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_label(batch_key))


def get_graph_row(batch_key, layer_idx):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_row(batch_key, layer_idx))


def get_graph_col(batch_key, layer_idx):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_col(batch_key, layer_idx))


def get_graph_data(batch_key, layer_idx):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_data(batch_key, layer_idx))


def get_graph_csc(batch_key, layer_idx):
    """(indptr i32[num_dst+1], indices i32[num_edge], edge_ids i32[num_edge] | None) of one layer: the block in the
    form `create_unitgraph_from_csc` takes (3rdparty/dgl.patch:30-57).  edge_ids is None when the layer is already
    ordered by destination (khop0/khop2/hash-dedup/random walk): indices then aliases get_graph_row()."""
    indptr, indices, eids = c_lib.samgraph_torch_get_graph_csc(batch_key, layer_idx)
    return _from_dlpack(indptr), _from_dlpack(indices), (None if eids is None else _from_dlpack(eids))


def get_dataset_feat():
    return _from_dlpack(c_lib.samgraph_torch_get_dataset_feat())


def get_dataset_label():
    return _from_dlpack(c_lib.samgraph_torch_get_dataset_label())


def get_graph_input_nodes(batch_key):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_input_nodes(batch_key))


def get_graph_output_nodes(batch_key):
    return _from_dlpack(c_lib.samgraph_torch_get_graph_output_nodes(batch_key))


def _create_dgl_block(data, num_src_nodes, num_dst_nodes):
    import dgl
    from dgl.heterograph import DGLBlock
    row, col = data
    gidx = dgl.heterograph_index.create_unitgraph_from_coo(2, num_src_nodes, num_dst_nodes, row, col, "coo")
    return DGLBlock(gidx, (["_N"], ["_N"]), ["_E"])


def _batch_tensors(batch_key, with_feat):
    if not with_feat:
        return None, None
    return get_graph_feat(batch_key), get_graph_label(batch_key)


def get_dgl_blocks(batch_key, num_layers, with_feat=True):
    feat, label = _batch_tensors(batch_key, with_feat)
    blocks = [_create_dgl_block((get_graph_row(batch_key, i), get_graph_col(batch_key, i)),
                                get_graph_num_src(batch_key, i), get_graph_num_dst(batch_key, i))
              for i in range(num_layers)]
    return blocks, feat, label


def get_dgl_blocks_with_weights(batch_key, num_layers, with_feat=True):
    feat, label = _batch_tensors(batch_key, with_feat)
    blocks = []
    for i in range(num_layers):
        block = _create_dgl_block((get_graph_row(batch_key, i), get_graph_col(batch_key, i)),
                                  get_graph_num_src(batch_key, i), get_graph_num_dst(batch_key, i))
        block.edata["weights"] = get_graph_data(batch_key, i)
        blocks.append(block)
    return blocks, feat, label


def _create_dgl_block_csc(batch_key, layer_idx):
    import dgl
    from dgl.heterograph import DGLBlock
    num_src, num_dst = get_graph_num_src(batch_key, layer_idx), get_graph_num_dst(batch_key, layer_idx)
    indptr, indices, eids = get_graph_csc(batch_key, layer_idx)
    if eids is None:
        eids = torch.arange(indices.shape[0], dtype=indices.dtype, device=indices.device)
    gidx = dgl.heterograph_index.create_unitgraph_from_csc(2, num_src, num_dst, indptr, indices, eids, "csc")
    return DGLBlock(gidx, (["_N"], ["_N"]), ["_E"])


def get_dgl_blocks_csc(batch_key, num_layers, with_feat=True):
    """get_dgl_blocks with the blocks created straight from CSC (needs the reference's DGL patch): DGL's own
    COO->CSC conversion — the reference's "convert" stage — disappears from the trainer."""
    feat, label = _batch_tensors(batch_key, with_feat)
    return [_create_dgl_block_csc(batch_key, i) for i in range(num_layers)], feat, label


def get_csc_blocks(batch_key, num_layers, with_feat=True):
    """DGL-free CSC hand-off: [(indptr, indices, edge_ids | None, num_src, num_dst)] per layer (e.g. for
    torch.sparse_csr_tensor / PyG SparseTensor(rowptr=indptr, col=indices))."""
    feat, label = _batch_tensors(batch_key, with_feat)
    blocks = [get_graph_csc(batch_key, i) + (get_graph_num_src(batch_key, i), get_graph_num_dst(batch_key, i))
              for i in range(num_layers)]
    return blocks, feat, label


def get_coo_blocks(batch_key, num_layers, with_feat=True):
    """DGL-free variant of get_dgl_blocks: [(row, col, num_src, num_dst)] per layer."""
    feat, label = _batch_tensors(batch_key, with_feat)
    blocks = [(get_graph_row(batch_key, i), get_graph_col(batch_key, i), get_graph_num_src(batch_key, i),
               get_graph_num_dst(batch_key, i)) for i in range(num_layers)]
    return blocks, feat, label


def notify_sampler_ready(barrier):
    barrier.wait()


def wait_for_sampler_ready(barrier):
    barrier.wait()


def load_subtensor(batch_key, feat, label, device):
    input_nodes = get_graph_input_nodes(batch_key).to(feat.device)
    output_nodes = get_graph_output_nodes(batch_key).to(label.device)
    batch_inputs = torch.index_select(feat, 0, input_nodes.long()).to(device)
    batch_labels = torch.index_select(label, 0, output_nodes.long()).to(device)
    return batch_inputs, batch_labels
