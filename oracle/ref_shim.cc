// TEST INFRASTRUCTURE — not product code.
//
// Thin shim that lets the reference's *unmodified* CPU hot-path translation
// units (compiled in place from /root/reference by oracle/Makefile, never
// copied) link without the reference's device.cc / cpu_device.cc, which call
// cudaHostAlloc and therefore need a GPU.  It provides
//   * a malloc-backed samgraph::common::Device (interface: device.h:36-62)
//   * a flat extern "C" surface over CPUSampleKHop0/2, CPUHashTable0/2 and
//     CPUExtract so Python (ctypes) can drive them.
// The result is oracle/_ref/libsamgraph_ref_cpu.so; it is used ONLY by
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm.
#include <omp.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "samgraph/common/common.h"
#include "samgraph/common/cpu/cpu_function.h"
#include "samgraph/common/cpu/cpu_hashtable.h"
#include "samgraph/common/cpu/cpu_hashtable0.h"
#include "samgraph/common/cpu/cpu_hashtable2.h"
#include "samgraph/common/device.h"
#include "samgraph/common/run_config.h"

namespace samgraph {
namespace common {

namespace {
class MallocDevice final : public Device {
 public:
  void SetDevice(Context) override {}
  void *AllocDataSpace(Context, size_t nbytes, size_t alignment) override {
    void *p = nullptr;
    if (nbytes == 0) nbytes = alignment;
    if (posix_memalign(&p, alignment, nbytes) != 0) abort();
    return p;
  }
  void FreeDataSpace(Context, void *ptr) override { free(ptr); }
  void CopyDataFromTo(const void *from, size_t from_offset, void *to,
                      size_t to_offset, size_t nbytes, Context, Context,
                      StreamHandle) override {
    memcpy(static_cast<char *>(to) + to_offset,
           static_cast<const char *>(from) + from_offset, nbytes);
  }
  void StreamSync(Context, StreamHandle) override {}
};
}  // namespace

// Base-class defaults normally supplied by device.cc.
void *Device::AllocWorkspace(Context ctx, size_t nbytes, double) {
  return AllocDataSpace(ctx, nbytes, kTempAllocaAlignment);
}
void Device::FreeWorkspace(Context ctx, void *ptr, size_t) {
  FreeDataSpace(ctx, ptr);
}
StreamHandle Device::CreateStream(Context) { return nullptr; }
void Device::FreeStream(Context, StreamHandle) {}
void Device::SyncStreamFromTo(Context, StreamHandle, StreamHandle) {}
Device *Device::Get(Context) {
  static MallocDevice dev;
  return &dev;
}

}  // namespace common
}  // namespace samgraph

using samgraph::common::IdType;
namespace sc = samgraph::common;

extern "C" {

void ref_set_omp_threads(int n) { sc::RunConfig::omp_thread_num = n; }
int ref_get_omp_threads() { return sc::RunConfig::omp_thread_num; }

// cpu_sampling_khop0.cc:29-83
void ref_cpu_sample_khop0(const uint32_t *indptr, const uint32_t *indices,
                          const uint32_t *input, size_t num_input,
                          uint32_t *out_src, uint32_t *out_dst,
                          size_t *num_out, size_t fanout) {
  sc::cpu::CPUSampleKHop0(indptr, indices, input, num_input, out_src, out_dst,
                          num_out, fanout);
}

// cpu_sampling_khop2.cc:29-76 (mutates indices in place)
void ref_cpu_sample_khop2(const uint32_t *indptr, uint32_t *indices,
                          const uint32_t *input, size_t num_input,
                          uint32_t *out_src, uint32_t *out_dst,
                          size_t *num_out, size_t fanout) {
  sc::cpu::CPUSampleKHop2(indptr, indices, input, num_input, out_src, out_dst,
                          num_out, fanout);
}

// kind 0 -> CPUHashTable0 (cpu_hashtable0.cc), 2 -> CPUHashTable2
void *ref_hashtable_new(int kind, size_t max_items) {
  sc::cpu::CPUHashTable *t = nullptr;
  if (kind == 0) t = new sc::cpu::CPUHashTable0(max_items);
  if (kind == 2) t = new sc::cpu::CPUHashTable2(max_items);
  return t;
}
void ref_hashtable_free(void *t) {
  delete static_cast<sc::cpu::CPUHashTable *>(t);
}
void ref_hashtable_populate(void *t, const uint32_t *input, size_t n) {
  static_cast<sc::cpu::CPUHashTable *>(t)->Populate(input, n);
}
void ref_hashtable_map_nodes(void *t, uint32_t *out, size_t n) {
  static_cast<sc::cpu::CPUHashTable *>(t)->MapNodes(out, n);
}
void ref_hashtable_map_edges(void *t, const uint32_t *src, const uint32_t *dst,
                             size_t len, uint32_t *new_src,
                             uint32_t *new_dst) {
  static_cast<sc::cpu::CPUHashTable *>(t)->MapEdges(src, dst, len, new_src,
                                                    new_dst);
}
void ref_hashtable_reset(void *t) {
  static_cast<sc::cpu::CPUHashTable *>(t)->Reset();
}
size_t ref_hashtable_num_items(void *t) {
  return static_cast<sc::cpu::CPUHashTable *>(t)->NumItems();
}

// cpu_extraction.cc:66-90 ; dtype follows common.h:38-46
void ref_cpu_extract(void *dst, const void *src, const uint32_t *index,
                     size_t num_index, size_t dim, int dtype) {
  sc::cpu::CPUExtract(dst, src, index, num_index, dim,
                      static_cast<sc::DataType>(dtype));
}

// common.cc:330-339
size_t ref_predict_num_nodes(size_t batch_size, const size_t *fanout,
                             size_t num_fanout) {
  std::vector<size_t> f(fanout, fanout + num_fanout);
  return sc::PredictNumNodes(batch_size, f, num_fanout);
}

}  // extern "C"
