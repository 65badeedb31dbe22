// Host runtime of the samgraph-compatible engine: shared types, logging, config.
//
// This mirrors the *interface* of the reference's samgraph/common (common.h,
// run_config.h, logging.h, profiler.h) so that the samgraph_* C-ABI and the
// Python API keep their meaning, while the implementation underneath is the
// device-count / single-sync design of the fgnn_k_* kernel layer.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "fgnn_kernels.h"

namespace fgnn {
namespace rt {

using IdType = uint32_t;  // common.h:35

// ---- enums: numeric values are part of the Python API (samgraph/common/__init__.py) ----
enum DataType { kF32 = 0, kF64 = 1, kF16 = 2, kU8 = 3, kI32 = 4, kI8 = 5, kI64 = 6 };  // common.h:38-46
enum DeviceType { kCPU = 0, kMMAP = 1, kGPU = 2 };                                        // common.h:48
enum SampleType {  // common.h:50-58
  kKHop0 = 0, kKHop1, kWeightedKHop, kRandomWalk, kWeightedKHopPrefix, kKHop2, kWeightedKHopHashDedup
};
enum RunArch { kArch0 = 0, kArch1, kArch2, kArch3, kArch4, kArch5, kArch6, kArch7 };  // common.h:69-78
enum CachePolicy {  // common.h:83-92
  kCacheByDegree = 0, kCacheByHeuristic, kCacheByPreSample, kCacheByDegreeHop, kCacheByPreSampleStatic,
  kCacheByFakeOptimal, kDynamicCache, kCacheByRandom
};

size_t DataTypeBytes(DataType t);

struct Context {  // common.h:94-106 ("cpu:0" / "cuda:1")
  DeviceType device_type = kCPU;
  int device_id = 0;
  Context() {}
  Context(DeviceType t, int id) : device_type(t), device_id(id) {}
  explicit Context(const std::string &name);
};

// ---- logging (logging.h:32-77): CHECK failures log and abort(), there are no error codes ----
enum LogLevel { kLogTrace = 0, kLogDebug, kLogInfo, kLogWarning, kLogError, kLogFatal };
LogLevel MinLogLevel();
class LogMessage : public std::ostringstream {
 public:
  LogMessage(const char *file, int line, LogLevel level, bool fatal) : file_(file), line_(line), level_(level), fatal_(fatal) {}
  ~LogMessage();
 private:
  const char *file_;
  int line_;
  LogLevel level_;
  bool fatal_;
};
#define FLOG(level) \
  if (::fgnn::rt::kLog##level >= ::fgnn::rt::MinLogLevel()) ::fgnn::rt::LogMessage(__FILE__, __LINE__, ::fgnn::rt::kLog##level, false)
#define FCHECK(x) \
  if (!(x)) ::fgnn::rt::LogMessage(__FILE__, __LINE__, ::fgnn::rt::kLogFatal, true) << "Check failed: " #x << ' '
#define FCHECK_EQ(a, b) FCHECK((a) == (b)) << "(" << (a) << " vs " << (b) << ") "
#define FCHECK_LE(a, b) FCHECK((a) <= (b)) << "(" << (a) << " vs " << (b) << ") "
#define FCHECK_LT(a, b) FCHECK((a) < (b)) << "(" << (a) << " vs " << (b) << ") "
#define CUDA_CALL(expr)                                                                        \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    FCHECK(e__ == cudaSuccess) << "CUDA: " << cudaGetErrorString(e__) << " in " #expr;         \
  } while (0)
#define FGNN_CALL(expr)                                                                        \
  do {                                                                                         \
    int e__ = (expr);                                                                          \
    FCHECK(e__ == 0) << "kernel ABI: " << fgnn_k_error_string(e__) << " in " #expr;            \
  } while (0)

class Timer {  // timer.h
 public:
  Timer() : t_(std::chrono::steady_clock::now()) {}
  double Passed() const {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_).count();
  }
  static uint64_t NowMicro() {
    return std::chrono::duration_cast<std::chrono::microseconds>(
               std::chrono::system_clock::now().time_since_epoch()).count();
  }
 private:
  std::chrono::steady_clock::time_point t_;
};

// ---- RunConfig (run_config.h:31-94; keys parsed as in operation.cc:45-169) ----
struct RunConfig {
  std::unordered_map<std::string, std::string> raw;
  std::string dataset_path;
  RunArch run_arch = kArch3;
  SampleType sample_type = kKHop2;
  size_t batch_size = 0, num_epoch = 0;
  Context sampler_ctx, trainer_ctx;
  CachePolicy cache_policy = kCacheByPreSample;
  double cache_percentage = 0.0;
  size_t max_sampling_jobs = 10, max_copying_jobs = 10;
  std::vector<size_t> fanout;
  size_t random_walk_length = 0, num_random_walk = 0, num_neighbor = 0, num_layer = 0;
  double random_walk_restart_prob = 0.0;
  bool is_configured = false;
  size_t num_sample_worker = 1, num_train_worker = 1;
  int barriered_epoch = 0, presample_epoch = 0;
  int omp_thread_num = 40;
  // environment (constant.cc:52-59)
  bool option_profile_cuda = false, option_sanity_check = false, option_dump_trace = false;
  size_t option_empty_feat = 0;
  // ours (optional keys / env): RNG seed, partitioned cache over trainer GPUs
  uint64_t seed = 0x46474E4E;
  bool partition_cache = false;
  // partitioned cache, hybrid layout: every trainer keeps the hottest replicate_percentage * V ranks, only the
  // rest of the cache is striped over the trainers (config key "replicate_percentage", env FGNN_REPLICATE_PCT)
  double replicate_percentage = 0.75;  // measured on 4 x B200: cold peer rows are expensive (profiles/r2_partition_diag_n4.txt)

  bool UseGPUCache() const { return cache_percentage > 0 && run_arch != kArch1; }  // run_config.h:81-83
  void LoadFromEnv();
  static RunConfig &Get();
};

size_t PredictNumNodes(size_t batch_size, const std::vector<size_t> &fanout, size_t upto);  // common.cc:330-339
std::string GetEnv(const std::string &k);
bool IsEnvSet(const std::string &k);

// ---- Tensor (common.h:108-160): data + dtype + shape + where it lives; freed when the last
// owner (engine or a Python tensor created from it) goes away. ----
class Tensor {
 public:
  ~Tensor();
  void *data = nullptr;
  DataType dtype = kI32;
  Context ctx;
  size_t nbytes = 0;
  std::vector<size_t> shape;
  std::string name;

  static std::shared_ptr<Tensor> Device(DataType dt, std::vector<size_t> shape, int device, cudaStream_t stream,
                                        const std::string &name);
  static std::shared_ptr<Tensor> Pinned(DataType dt, std::vector<size_t> shape, const std::string &name);
  static std::shared_ptr<Tensor> FromMmap(const std::string &path, DataType dt, std::vector<size_t> shape,
                                          const std::string &name);
  static std::shared_ptr<Tensor> View(void *data, DataType dt, std::vector<size_t> shape, Context ctx,
                                      std::shared_ptr<void> keep_alive, const std::string &name);
  size_t NumItems() const;

 private:
  enum Kind { kNone, kDeviceAsync, kPinned, kMmap, kView } kind_ = kNone;
  cudaStream_t stream_ = nullptr;
  std::shared_ptr<void> keep_;
};
using TensorPtr = std::shared_ptr<Tensor>;

struct TrainGraph {  // common.h:178-186
  TensorPtr row, col, data;
  size_t num_src = 0, num_dst = 0, num_edge = 0;
  // CSC form of the same block, built on first request by the hand-off (pymodule.cc GetGraphCsc):
  // indptr u32[num_dst+1], indices u32[num_edge] (aliases `row` when col is already ascending), edge ids
  TensorPtr csc_indptr, csc_indices, csc_eids;
};

struct Task {  // common.h:197-217
  uint64_t key = 0;
  std::vector<TrainGraph> graphs;
  TensorPtr input_nodes, output_nodes, input_feat, output_label;
  size_t num_miss = 0, num_cache = 0;
  int slot = -1;                // sampler slot (stream + scratch) that produces this batch
  std::shared_ptr<void> block;  // pooled device block the sampler outputs live in (rt_engine.cc TaskBlock)
  cudaEvent_t ready = nullptr;  // recorded on the producing stream when all tensors are final
  cudaEvent_t extracted = nullptr;  // recorded on the extraction stream after feature + label gather
  int xslot = 0;
  cudaEvent_t xbegin = nullptr;  // FGNN_TRACE_GPU=1
  uint64_t t_extract = 0;       // host clock (us) when extraction was enqueued
  ~Task();
};
using TaskPtr = std::shared_ptr<Task>;

struct Dataset {  // common.h:128-160
  TensorPtr indptr, indices, prob_table, alias_table, prob_prefix_table, feat, label, train_set, test_set, valid_set,
      ranking_nodes;
  size_t num_node = 0, num_edge = 0, num_class = 0, feat_dim = 0;
};

}  // namespace rt
}  // namespace fgnn
