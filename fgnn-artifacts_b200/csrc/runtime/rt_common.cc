#include "rt_common.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <ctime>
#include <map>
#include <mutex>
#include <numeric>

namespace fgnn {
namespace rt {

size_t DataTypeBytes(DataType t) {  // common.cc GetDataTypeBytes
  switch (t) {
    case kI8: case kU8: return 1;
    case kF16: return 2;
    case kF32: case kI32: return 4;
    case kI64: case kF64: return 8;
  }
  FCHECK(false) << "bad dtype";
  return 0;
}

Context::Context(const std::string &name) {  // common.cc Context(std::string): "cpu:0", "cuda:1", "mmap:0"
  const size_t sep = name.find(':');
  FCHECK(sep != std::string::npos) << "bad context string " << name;
  std::string dev = name.substr(0, sep);
  std::transform(dev.begin(), dev.end(), dev.begin(), ::tolower);
  device_id = std::stoi(name.substr(sep + 1));
  if (dev == "cpu") device_type = kCPU;
  else if (dev == "cuda" || dev == "gpu") device_type = kGPU;
  else if (dev == "mmap") device_type = kMMAP;
  else FCHECK(false) << "bad context string " << name;
}

std::string GetEnv(const std::string &k) {
  const char *v = getenv(k.c_str());
  return v ? std::string(v) : std::string();
}
bool IsEnvSet(const std::string &k) {  // common.cc IsEnvSet
  const std::string v = GetEnv(k);
  return v == "1" || v == "ON" || v == "On" || v == "on";
}

LogLevel MinLogLevel() {  // logging.cc:96-110 SAMGRAPH_LOG_LEVEL
  static int level = -1;
  if (level < 0) {
    std::string v = GetEnv("SAMGRAPH_LOG_LEVEL");
    std::transform(v.begin(), v.end(), v.begin(), ::tolower);
    if (v == "trace") level = kLogTrace;
    else if (v == "debug") level = kLogDebug;
    else if (v == "info") level = kLogInfo;
    else if (v == "warn" || v == "warning") level = kLogWarning;
    else if (v == "error") level = kLogError;
    else if (v == "fatal") level = kLogFatal;
    else level = kLogWarning;
  }
  return (LogLevel)level;
}

LogMessage::~LogMessage() {
  static const char *names[] = {"TRACE", "DEBUG", "INFO", "WARNING", "ERROR", "FATAL"};
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lk(mu);
    if (!IsEnvSet("SAMGRAPH_LOG_HIDE_TIME")) {
      char buf[32];
      time_t t = time(nullptr);
      struct tm tmv;
      localtime_r(&t, &tmv);
      strftime(buf, sizeof(buf), "%F %T", &tmv);
      fprintf(stderr, "[%s] ", buf);
    }
    fprintf(stderr, "%s:%d: %s: %s\n", file_, line_, names[level_], str().c_str());
    fflush(stderr);
  }
  if (fatal_) abort();  // logging.cc:69-73
}

RunConfig &RunConfig::Get() {
  static RunConfig rc;
  return rc;
}

void RunConfig::LoadFromEnv() {  // run_config.cc:78-101
  option_profile_cuda = IsEnvSet("SAMGRAPH_PROFILE_CUDA");
  option_sanity_check = IsEnvSet("SAMGRAPH_SANITY_CHECK");
  option_dump_trace = IsEnvSet("SAMGRAPH_DUMP_TRACE");
  if (!GetEnv("SAMGRAPH_EMPTY_FEAT").empty()) option_empty_feat = std::stoul(GetEnv("SAMGRAPH_EMPTY_FEAT"));
  if (!GetEnv("FGNN_SEED").empty()) seed = std::stoull(GetEnv("FGNN_SEED"));
  if (IsEnvSet("FGNN_PARTITION_CACHE")) partition_cache = true;
}

size_t PredictNumNodes(size_t batch_size, const std::vector<size_t> &fanout, size_t upto) {
  FCHECK_LE(upto, fanout.size());
  size_t count = batch_size;
  for (int i = (int)upto - 1; i >= 0; --i) count += count * fanout[i];
  return count;
}

// ---------------------------------------------------------------------------
size_t Tensor::NumItems() const {
  return std::accumulate(shape.begin(), shape.end(), (size_t)1, std::multiplies<size_t>());
}

// ---------------------------------------------------------------------------
// WorkspacePool (workspace_pool.cc:30-196, cuda_device.cc:143-157): size-sorted free list per GPU, 4 KiB
// rounding, blocks over-allocated x1.25 the first time a size is seen (constant.h:78) so the slightly
// different batch sizes of consecutive mini-batches reuse the same blocks.  Backing store is plain
// cudaMalloc; nothing is returned to the driver until the process ends (or an allocation fails).
// The stream-ordered allocator (cudaMallocAsync) was measured unusable here: 283 MB feature tensors whose
// size changes every batch made it remap on most calls — 1-8 ms per allocation, occasionally 100+ ms
// (gpurun r1_k) — which stalled the pump while the GPU sat idle.
// Memory is reusable as soon as it is freed: callers free only after the producing/consuming GPU work has
// been synchronised on the host (Sampler::Finish / Extractor::Finish), like the reference.
// ---------------------------------------------------------------------------
namespace {
class DevicePool {
 public:
  void *Alloc(int device, size_t nbytes) {
    const size_t want = RoundUp(std::max<size_t>(nbytes, 256));
    {
      std::lock_guard<std::mutex> lk(mu_);
      auto &fl = free_[device];
      auto it = fl.lower_bound(want);
      // smallest block that fits, unless it would waste more than half of a big block
      if (it != fl.end() && (it->first <= want * 2 || it->first <= (1u << 20))) {
        void *p = it->second;
        fl.erase(it);
        return p;
      }
    }
    const size_t cap = want >= (1u << 20) ? RoundUp(want + want / 4) : want;
    int cur = 0;
    CUDA_CALL(cudaGetDevice(&cur));
    if (cur != device) CUDA_CALL(cudaSetDevice(device));
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
      cudaGetLastError();
      Trim(device);
      e = cudaMalloc(&p, cap);
    }
    FCHECK(e == cudaSuccess) << "cudaMalloc(" << cap << " B) on GPU " << device << ": " << cudaGetErrorString(e);
    if (cur != device) CUDA_CALL(cudaSetDevice(cur));
    std::lock_guard<std::mutex> lk(mu_);
    size_[p] = cap;
    return p;
  }
  void Free(int device, void *p) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = size_.find(p);
    if (it == size_.end()) return;
    free_[device].emplace(it->second, p);
  }
  void Trim(int device) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto &kv : free_[device]) {
      cudaFree(kv.second);
      size_.erase(kv.second);
    }
    free_[device].clear();
    cudaGetLastError();
  }
  static DevicePool &Get() {
    static DevicePool *p = new DevicePool();  // leaked on purpose: tensors may outlive static destructors
    return *p;
  }

 private:
  static size_t RoundUp(size_t n) { return (n + 4095) & ~(size_t)4095; }
  std::mutex mu_;
  std::map<int, std::multimap<size_t, void *>> free_;
  std::unordered_map<void *, size_t> size_;
};
}  // namespace

TensorPtr Tensor::Device(DataType dt, std::vector<size_t> shape, int device, cudaStream_t stream,
                         const std::string &name) {
  auto t = std::make_shared<Tensor>();
  t->dtype = dt;
  t->shape = std::move(shape);
  t->ctx = Context(kGPU, device);
  t->nbytes = t->NumItems() * DataTypeBytes(dt);
  t->name = name;
  t->kind_ = kDeviceAsync;
  t->stream_ = stream;
  t->data = DevicePool::Get().Alloc(device, t->nbytes);
  return t;
}

TensorPtr Tensor::Pinned(DataType dt, std::vector<size_t> shape, const std::string &name) {
  auto t = std::make_shared<Tensor>();
  t->dtype = dt;
  t->shape = std::move(shape);
  t->ctx = Context(kCPU, 0);
  t->nbytes = t->NumItems() * DataTypeBytes(dt);
  t->name = name;
  t->kind_ = kPinned;
  CUDA_CALL(cudaHostAlloc(&t->data, std::max<size_t>(t->nbytes, 16), cudaHostAllocMapped | cudaHostAllocPortable));
  return t;
}

TensorPtr Tensor::FromMmap(const std::string &path, DataType dt, std::vector<size_t> shape,
                           const std::string &name) {  // common.cc Tensor::FromMmap
  auto t = std::make_shared<Tensor>();
  t->dtype = dt;
  t->shape = std::move(shape);
  t->ctx = Context(kMMAP, 0);
  t->nbytes = t->NumItems() * DataTypeBytes(dt);
  t->name = name;
  t->kind_ = kMmap;
  struct stat st;
  FCHECK(stat(path.c_str(), &st) == 0) << "cannot stat " << path;
  FCHECK_LE(t->nbytes, (size_t)st.st_size) << path << " is smaller than its meta.txt shape";
  int fd = open(path.c_str(), O_RDONLY);
  FCHECK(fd >= 0) << "cannot open " << path;
  if (t->nbytes > 0) {
    t->data = mmap(nullptr, t->nbytes, PROT_READ, MAP_SHARED, fd, 0);
    FCHECK(t->data != MAP_FAILED) << "mmap failed for " << path;
  }
  close(fd);
  return t;
}

TensorPtr Tensor::View(void *data, DataType dt, std::vector<size_t> shape, Context ctx,
                       std::shared_ptr<void> keep_alive, const std::string &name) {
  auto t = std::make_shared<Tensor>();
  t->data = data;
  t->dtype = dt;
  t->shape = std::move(shape);
  t->ctx = ctx;
  t->nbytes = t->NumItems() * DataTypeBytes(dt);
  t->name = name;
  t->kind_ = kView;
  t->keep_ = std::move(keep_alive);
  return t;
}

Tensor::~Tensor() {
  if (!data) return;
  switch (kind_) {
    case kDeviceAsync:
      // may run on a Python thread, possibly during interpreter teardown: never abort here
      DevicePool::Get().Free(ctx.device_id, data);
      break;
    case kPinned:
      cudaFreeHost(data);
      cudaGetLastError();
      break;
    case kMmap:
      munmap(data, nbytes);
      break;
    default:
      break;
  }
}

Task::~Task() {
  if (ready) cudaEventDestroy(ready);
  if (extracted) cudaEventDestroy(extracted);
  if (xbegin) cudaEventDestroy(xbegin);
  cudaGetLastError();
}

}  // namespace rt
}  // namespace fgnn
