// Python module `samgraph.torch.c_lib`: the tensor hand-off of the reference's pybind adapter
// (samgraph/torch/adapter.cc:48-192) without compiling against torch.  Every getter returns a
// DLPack capsule that aliases the engine's buffer (zero copy); samgraph/torch/adapter.py turns it
// into a torch.Tensor with torch.from_dlpack.  Ownership follows adapter.cc:54-62: the capsule
// keeps the engine Tensor alive until Python drops the torch tensor, then the memory returns to
// the stream-ordered pool.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include "rt_engine.h"

using namespace fgnn::rt;

namespace {

// ---- DLPack v0.8 ABI (dlpack.h), declared locally: the structs are a stable C ABI ----
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
  void *data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t *shape;
  int64_t *strides;
  uint64_t byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void *manager_ctx;
  void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;

struct Holder {
  TensorPtr tensor;
  TaskPtr task;  // keeps the `ready` event and sibling tensors valid
  int64_t shape[4];
};

void ManagedDeleter(DLManagedTensor *self) {
  delete static_cast<Holder *>(self->manager_ctx);
  delete self;
}

void CapsuleDestructor(PyObject *cap) {
  // consumed capsules are renamed "used_dltensor" by the importer and must not be freed here
  if (!PyCapsule_IsValid(cap, "dltensor")) return;
  auto *m = static_cast<DLManagedTensor *>(PyCapsule_GetPointer(cap, "dltensor"));
  if (m && m->deleter) m->deleter(m);
}

PyObject *ToCapsule(const TensorPtr &t, const TaskPtr &task, DataType view_as) {
  if (!t || (!t->data && t->nbytes)) {
    PyErr_SetString(PyExc_RuntimeError, "samgraph: tensor is not available for this batch");
    return nullptr;
  }
  auto *h = new Holder();
  h->tensor = t;
  h->task = task;
  auto *m = new DLManagedTensor();
  m->dl_tensor.data = t->data;
  m->dl_tensor.device.device_type = (t->ctx.device_type == kGPU) ? kDLCUDA : kDLCPU;
  m->dl_tensor.device.device_id = (t->ctx.device_type == kGPU) ? t->ctx.device_id : 0;
  m->dl_tensor.ndim = (int32_t)t->shape.size();
  for (size_t i = 0; i < t->shape.size() && i < 4; ++i) h->shape[i] = (int64_t)t->shape[i];
  m->dl_tensor.shape = h->shape;
  m->dl_tensor.strides = nullptr;
  m->dl_tensor.byte_offset = 0;
  DLDataType dt;
  dt.lanes = 1;
  switch (view_as) {
    case kF32: dt.code = 2; dt.bits = 32; break;
    case kF64: dt.code = 2; dt.bits = 64; break;
    case kF16: dt.code = 2; dt.bits = 16; break;
    case kU8: dt.code = 1; dt.bits = 8; break;
    case kI8: dt.code = 0; dt.bits = 8; break;
    case kI32: dt.code = 0; dt.bits = 32; break;   // ids are exposed as int32, like adapter.cc (kI32)
    case kI64: dt.code = 0; dt.bits = 64; break;
  }
  m->dl_tensor.dtype = dt;
  m->manager_ctx = h;
  m->deleter = ManagedDeleter;
  return PyCapsule_New(m, "dltensor", CapsuleDestructor);
}

// torch.Tensor view of the same memory (adapter.cc:54-62 returns torch::from_blob tensors).  The extension does not
// link libtorch: the DLPack capsule is adopted by torch.utils.dlpack.from_dlpack, looked up once through the CPython
// API, so `samgraph_torch_*` hand out real tensors and the reference's adapter.py binds unmodified.
PyObject *ToTensor(const TensorPtr &t, const TaskPtr &task, DataType view_as) {
  static PyObject *from_dlpack = nullptr;
  if (!from_dlpack) {
    PyObject *mod = PyImport_ImportModule("torch.utils.dlpack");
    if (!mod) return nullptr;
    from_dlpack = PyObject_GetAttrString(mod, "from_dlpack");
    Py_DECREF(mod);
    if (!from_dlpack) return nullptr;
  }
  PyObject *cap = ToCapsule(t, task, view_as);
  if (!cap) return nullptr;
  PyObject *tensor = PyObject_CallFunctionObjArgs(from_dlpack, cap, nullptr);
  Py_DECREF(cap);
  return tensor;
}

TaskPtr Batch(unsigned long long key) {
  TaskPtr b = Engine::Get()->CurrentBatch();
  if (!b) {
    PyErr_SetString(PyExc_RuntimeError, "samgraph: no current batch (call get_next_batch first)");
    return nullptr;
  }
  FCHECK_EQ((uint64_t)key, b->key);  // adapter.cc:54
  return b;
}

PyObject *GetGraphFeat(PyObject *, PyObject *args) {
  unsigned long long key;
  if (!PyArg_ParseTuple(args, "K", &key)) return nullptr;
  TaskPtr b = Batch(key);
  return b ? ToTensor(b->input_feat, b, kF32) : nullptr;
}
PyObject *GetGraphLabel(PyObject *, PyObject *args) {
  unsigned long long key;
  if (!PyArg_ParseTuple(args, "K", &key)) return nullptr;
  TaskPtr b = Batch(key);
  return b ? ToTensor(b->output_label, b, kI64) : nullptr;
}
template <int WHICH>
PyObject *GetGraphEdge(PyObject *, PyObject *args) {
  unsigned long long key;
  int layer;
  if (!PyArg_ParseTuple(args, "Ki", &key, &layer)) return nullptr;
  TaskPtr b = Batch(key);
  if (!b) return nullptr;
  if (layer < 0 || layer >= (int)b->graphs.size()) {
    PyErr_SetString(PyExc_IndexError, "samgraph: layer index out of range");
    return nullptr;
  }
  const TrainGraph &g = b->graphs[layer];
  return ToTensor(WHICH == 0 ? g.row : (WHICH == 1 ? g.col : g.data), b, kI32);
}
// Block hand-off in CSC form: (indptr, indices, edge_ids | None) of one layer, the arguments of
// the reference's DGL patch `create_unitgraph_from_csc` (3rdparty/dgl.patch:30-57).  The reference builds COO
// blocks (adapter.py:92-95) and pays DGL's COO->CSC conversion in the trainer (kLogL1ConvertTime).
// khop0 / khop2 / hash-dedup / random-walk layers are emitted seed-major, so `indices` IS `row` (zero copy,
// edge_ids = None = identity) and only the indptr is computed; khop1 / weighted layers are sorted by dst.
cudaStream_t HandoffStream(int device) {
  static cudaStream_t streams[64] = {nullptr};
  FCHECK(device >= 0 && device < 64);
  if (!streams[device]) CUDA_CALL(cudaStreamCreateWithFlags(&streams[device], cudaStreamNonBlocking));
  return streams[device];
}
PyObject *GetGraphCsc(PyObject *, PyObject *args) {
  unsigned long long key;
  int layer;
  if (!PyArg_ParseTuple(args, "Ki", &key, &layer)) return nullptr;
  TaskPtr b = Batch(key);
  if (!b) return nullptr;
  if (layer < 0 || layer >= (int)b->graphs.size()) {
    PyErr_SetString(PyExc_IndexError, "samgraph: layer index out of range");
    return nullptr;
  }
  TrainGraph &g = b->graphs[layer];
  if (!g.row || !g.col || g.row->ctx.device_type != kGPU) {
    PyErr_SetString(PyExc_RuntimeError, "samgraph: the block's row/col are not on a GPU");
    return nullptr;
  }
  const SampleType st = RunConfig::Get().sample_type;
  const bool sorted = st == kKHop0 || st == kKHop2 || st == kWeightedKHopHashDedup || st == kRandomWalk ||
                      g.num_edge == 0;
  if (!g.csc_indptr) {
    const int dev = g.row->ctx.device_id;
    int cur = 0;
    CUDA_CALL(cudaGetDevice(&cur));
    if (cur != dev) CUDA_CALL(cudaSetDevice(dev));
    cudaStream_t s = HandoffStream(dev);
    const uint32_t e = (uint32_t)g.num_edge, nd = (uint32_t)g.num_dst;
    TensorPtr indptr = Tensor::Device(kI32, {g.num_dst + 1}, dev, s, "csc_indptr");
    TensorPtr indices, eids, ws;
    if (sorted) {
      indices = g.row;
      FGNN_CALL(fgnn_k_coo_to_csc((const uint32_t *)g.row->data, (const uint32_t *)g.col->data, e, nullptr, nd, 1,
                                  (uint32_t *)indptr->data, nullptr, nullptr, nullptr, 0, s));
    } else {
      indices = Tensor::Device(kI32, {g.num_edge}, dev, s, "csc_indices");
      eids = Tensor::Device(kI32, {g.num_edge}, dev, s, "csc_eids");
      const size_t wb = fgnn_k_coo_to_csc_workspace_bytes(e, nd);
      ws = Tensor::Device(kU8, {wb}, dev, s, "csc_ws");
      FGNN_CALL(fgnn_k_coo_to_csc((const uint32_t *)g.row->data, (const uint32_t *)g.col->data, e, nullptr, nd, 0,
                                  (uint32_t *)indptr->data, (uint32_t *)indices->data, (uint32_t *)eids->data,
                                  ws->data, wb, s));
    }
    CUDA_CALL(cudaStreamSynchronize(s));  // the pool is not stream-ordered: finish before `ws` is recycled
    if (cur != dev) CUDA_CALL(cudaSetDevice(cur));
    g.csc_indptr = indptr;
    g.csc_indices = indices;
    g.csc_eids = eids;
  }
  PyObject *a = ToTensor(g.csc_indptr, b, kI32);
  PyObject *c = a ? ToTensor(g.csc_indices, b, kI32) : nullptr;
  PyObject *d = nullptr;
  if (c) {
    if (g.csc_eids) d = ToTensor(g.csc_eids, b, kI32);
    else { d = Py_None; Py_INCREF(d); }
  }
  if (!a || !c || !d) {
    Py_XDECREF(a); Py_XDECREF(c); Py_XDECREF(d);
    return nullptr;
  }
  PyObject *t = PyTuple_Pack(3, a, c, d);
  Py_DECREF(a); Py_DECREF(c); Py_DECREF(d);
  return t;
}
PyObject *GetInputNodes(PyObject *, PyObject *args) {
  unsigned long long key;
  if (!PyArg_ParseTuple(args, "K", &key)) return nullptr;
  TaskPtr b = Batch(key);
  return b ? ToTensor(b->input_nodes, b, kI32) : nullptr;
}
PyObject *GetOutputNodes(PyObject *, PyObject *args) {
  unsigned long long key;
  if (!PyArg_ParseTuple(args, "K", &key)) return nullptr;
  TaskPtr b = Batch(key);
  return b ? ToTensor(b->output_nodes, b, kI32) : nullptr;
}
PyObject *GetDatasetFeat(PyObject *, PyObject *) {
  const Dataset *ds = Engine::Get()->GetDataset();
  if (!ds) { PyErr_SetString(PyExc_RuntimeError, "samgraph: dataset not loaded"); return nullptr; }
  return ToTensor(ds->feat, nullptr, kF32);
}
PyObject *GetDatasetLabel(PyObject *, PyObject *) {
  const Dataset *ds = Engine::Get()->GetDataset();
  if (!ds) { PyErr_SetString(PyExc_RuntimeError, "samgraph: dataset not loaded"); return nullptr; }
  return ToTensor(ds->label, nullptr, kI64);
}

PyMethodDef kMethods[] = {
    {"samgraph_torch_get_graph_feat", GetGraphFeat, METH_VARARGS, "torch.Tensor: f32 [num_input, feat_dim] on the trainer GPU"},
    {"samgraph_torch_get_graph_label", GetGraphLabel, METH_VARARGS, "torch.Tensor: i64 [batch] on the trainer GPU"},
    {"samgraph_torch_get_graph_row", GetGraphEdge<0>, METH_VARARGS, "torch.Tensor: i32 [num_edge] (neighbour local ids)"},
    {"samgraph_torch_get_graph_col", GetGraphEdge<1>, METH_VARARGS, "torch.Tensor: i32 [num_edge] (seed local ids)"},
    {"samgraph_torch_get_graph_data", GetGraphEdge<2>, METH_VARARGS, "torch.Tensor: i32 [num_edge] (random-walk visit counts)"},
    {"samgraph_torch_get_graph_csc", GetGraphCsc, METH_VARARGS, "(indptr i32 [num_dst+1], indices i32 [num_edge], edge_ids i32 [num_edge] | None) tensors"},
    {"samgraph_torch_get_dataset_feat", GetDatasetFeat, METH_NOARGS, "torch.Tensor: host feature table"},
    {"samgraph_torch_get_dataset_label", GetDatasetLabel, METH_NOARGS, "torch.Tensor: host label table"},
    {"samgraph_torch_get_graph_input_nodes", GetInputNodes, METH_VARARGS, "torch.Tensor: i32 [num_input]"},
    {"samgraph_torch_get_graph_output_nodes", GetOutputNodes, METH_VARARGS, "torch.Tensor: i32 [batch]"},
    {nullptr, nullptr, 0, nullptr}};

PyModuleDef kModule = {PyModuleDef_HEAD_INIT, "c_lib", "samgraph B200 runtime: tensor hand-off", -1, kMethods,
                       nullptr, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit_c_lib(void) { return PyModule_Create(&kModule); }
