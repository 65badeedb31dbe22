#include "rt_engine.h"

#include "fgnn_dataset_tools.h"

#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <iterator>

namespace fgnn {
namespace rt {

}  // namespace rt
}  // namespace fgnn

// Share of an epoch's steps owned by sampler `worker_id` of `num_worker` (dist_shuffler.cc:60-83):
// every sampler takes floor(N/S) consecutive steps, the last one also takes the remainder.
extern "C" void fgnn_rt_step_split(size_t num_step, size_t num_worker, size_t worker_id, size_t *begin,
                                   size_t *count) {
  const size_t per = num_step / num_worker;
  *begin = per * worker_id;
  *count = (worker_id == num_worker - 1) ? num_step - *begin : per;
}

namespace fgnn {
namespace rt {

// =============================================================================================
// TaskPool
// =============================================================================================
bool TaskPool::Full() {
  std::lock_guard<std::mutex> lk(mu_);
  return q_.size() >= max_;
}
void TaskPool::Submit(TaskPtr t) {
  std::lock_guard<std::mutex> lk(mu_);
  q_.push_back(std::move(t));
}
TaskPtr TaskPool::TryGet() {
  std::lock_guard<std::mutex> lk(mu_);
  if (q_.empty()) return nullptr;
  TaskPtr t = q_.front();
  q_.pop_front();
  return t;
}
TaskPtr TaskPool::Get(std::atomic<bool> *stop) {
  while (true) {
    TaskPtr t = TryGet();
    if (t) return t;
    if (stop && stop->load()) return nullptr;
    std::this_thread::sleep_for(std::chrono::microseconds(1));
  }
}

// =============================================================================================
// Shared ring (arch5): sampler processes -> trainer processes through pinned host memory.
// Record layout follows TransData/GraphData (task_queue.cc:68-88): a header, then
// [input_nodes][output_nodes] and per layer {num_src,num_dst,num_edge,row,col[,data]}.
// =============================================================================================
struct SlotHeader {
  std::atomic<uint32_t> ready;  // the slot's sequence word (rt_ring.h)
  uint32_t num_layer, have_data;
  uint64_t key;
  uint64_t input_size, output_size;
  uint64_t num_src[8], num_dst[8], num_edge[8];
};

struct SharedRing {
  RingCtl ctl;  // ticket protocol: rt_ring.h
  pthread_barrier_t sampler_barrier, trainer_barrier;
  uint64_t slot_bytes;
  uint64_t ranking_off, slots_off;
  std::atomic<uint32_t> presample_done;
  std::atomic<uint32_t> live_workers, left_workers;  // shutdown rendezvous of the arch5 worker processes
  // partitioned cache: one IPC handle per trainer
  unsigned char ipc_handle[16][FGNN_IPC_HANDLE_BYTES];
  uint64_t shard_rows[16];
  // device queue: the slot PAYLOADS live in trainer HBM, slot s on trainer s % T at local
  // index s / T; samplers write them with peer copies over NVLink, only the SlotHeader stays in host memory
  uint32_t devq_enabled, devq_trainers;
  std::atomic<uint32_t> devq_ready[16];
  unsigned char devq_handle[16][FGNN_IPC_HANDLE_BYTES];
  char *base() { return reinterpret_cast<char *>(this); }
  IdType *ranking() { return reinterpret_cast<IdType *>(base() + ranking_off); }
  char *slot(uint64_t i) { return base() + slots_off + (i % ctl.num_slots) * slot_bytes; }
  std::atomic<uint32_t> *ready_of(uint64_t i);
};

std::atomic<uint32_t> *SharedRing::ready_of(uint64_t i) { return &reinterpret_cast<SlotHeader *>(slot(i))->ready; }

// =============================================================================================
// Sampler: DoShuffle + DoGPUSample (cuda_loops.cc:30-267 == dist_loops.cc:35-269)
// =============================================================================================
// One device allocation holding every sampler output of a mini-batch at its PredictNumNodes bound: the seed
// list, the running unique list (== input_nodes) and per layer row / col / data.  The kernels write straight
// into the block and the task's tensors are exact-size VIEWS of it, so the hot path has no allocation and no
// device-to-device "materialise" copies; 180 GB of HBM make a pool of bound-sized blocks cheap (27 MB each for
// GraphSAGE [25,10], 100 MB for GCN [5,10,15]).  A block returns to the pool when the last tensor that views
// it is released (Python included).
struct TaskBlock {
  char *base = nullptr;
  size_t bytes = 0;
  IdType *seeds = nullptr, *n2o = nullptr;
  IdType *row[8] = {nullptr}, *col[8] = {nullptr}, *data[8] = {nullptr};
};

class BlockPool : public std::enable_shared_from_this<BlockPool> {
 public:
  BlockPool(int dev, size_t batch, size_t max_nodes, const std::vector<size_t> &edge_max, bool with_data)
      : dev_(dev), batch_(batch), max_nodes_(max_nodes), edge_max_(edge_max), with_data_(with_data) {}
  ~BlockPool() {
    cudaSetDevice(dev_);
    for (TaskBlock *b : all_) { cudaFree(b->base); delete b; }
    cudaGetLastError();
  }
  void Reserve(size_t n) {
    std::lock_guard<std::mutex> lk(mu_);
    while (all_.size() < n) free_.push_back(NewBlock());
  }
  // never blocks: a consumer that holds on to many batches makes the pool grow instead of stalling the sampler
  std::shared_ptr<TaskBlock> Acquire() {
    TaskBlock *b = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (free_.empty()) free_.push_back(NewBlock());
      // FIFO: the block that has been free the longest.  A block returns to the pool when Python drops the last
      // tensor that views it; kernels the consumer queued on ITS stream may still read it (the reference's
      // WorkspacePool has the same hazard, common.cc:68-75), so the most recently freed block is reused last.
      b = free_.front();
      free_.pop_front();
    }
    std::shared_ptr<BlockPool> self = shared_from_this();
    return std::shared_ptr<TaskBlock>(b, [self](TaskBlock *x) {
      std::lock_guard<std::mutex> lk(self->mu_);
      self->free_.push_back(x);
    });
  }
  size_t NumBlocks() { std::lock_guard<std::mutex> lk(mu_); return all_.size(); }

 private:
  TaskBlock *NewBlock() {  // mu_ held
    auto al = [](size_t n) { return (n * sizeof(IdType) + 255) & ~(size_t)255; };
    size_t total = al(batch_) + al(max_nodes_ + 1);
    for (size_t e : edge_max_) total += al(e + 1) * (with_data_ ? 3 : 2);
    TaskBlock *b = new TaskBlock();
    int cur = 0;
    CUDA_CALL(cudaGetDevice(&cur));
    if (cur != dev_) CUDA_CALL(cudaSetDevice(dev_));
    CUDA_CALL(cudaMalloc((void **)&b->base, total));
    if (cur != dev_) CUDA_CALL(cudaSetDevice(cur));
    b->bytes = total;
    char *p = b->base;
    b->seeds = (IdType *)p; p += al(batch_);
    b->n2o = (IdType *)p; p += al(max_nodes_ + 1);
    for (size_t i = 0; i < edge_max_.size(); ++i) {
      b->row[i] = (IdType *)p; p += al(edge_max_[i] + 1);
      b->col[i] = (IdType *)p; p += al(edge_max_[i] + 1);
      if (with_data_) { b->data[i] = (IdType *)p; p += al(edge_max_[i] + 1); }
    }
    all_.push_back(b);
    return b;
  }
  int dev_;
  size_t batch_, max_nodes_;
  std::vector<size_t> edge_max_;
  bool with_data_;
  std::mutex mu_;
  std::vector<TaskBlock *> all_;
  std::deque<TaskBlock *> free_;
};

// One sampling slot = one mini-batch in flight: its own stream, hash table and scratch.  A single 8000-seed
// batch cannot fill a B200 (ncu r1_b: sampling kernels at 3-27 % active warps, every kernel a chain of
// dependent DRAM latencies), so the sampler keeps `num_slots` batches in flight on separate streams and
// hands them out in order.
// FGNN_TRACE_HOST=1: report any host-side region of the hot loop that blocks for more than 1 ms
struct SlowScope {
  const char *name;
  Timer t;
  explicit SlowScope(const char *n) : name(n) {}
  ~SlowScope() {
    static const bool on = IsEnvSet("FGNN_TRACE_HOST");
    if (on && t.Passed() > 1e-3) fprintf(stderr, "[fgnn slow] %s %.3f ms\n", name, t.Passed() * 1e3);
  }
};

static bool TraceGpu() {
  static const bool on = IsEnvSet("FGNN_TRACE_GPU");
  return on;
}

struct SampleSlot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;   // counts are on the host
  cudaEvent_t idle = nullptr;   // scratch marker used by Reshuffle
  cudaEvent_t begin = nullptr;  // FGNN_TRACE_GPU=1: device-side start of the batch
  TensorPtr table, num_items, chain, ws, rank_ws;
  uint32_t *counts_dev = nullptr, *counts_host = nullptr;  // views of the sampler's contiguous count arrays
  std::vector<TensorPtr> dst, pos;   // scratch: sampled global ids, their bucket positions
  uint32_t ver_state = 0;            // versioned table reset (fgnn_k_ht_next_version)
};

class Sampler {
 public:
  Sampler(const Dataset *ds, Context ctx, int worker_id, int num_worker, size_t num_epoch);
  ~Sampler();
  // next mini-batch of this sampler's share of the epoch; nullptr when all epochs are done
  TaskPtr Next();
  // Super-batch: the next (up to) GroupSize() mini-batches, on consecutive slots of one slot group, enqueued with
  // ONE call into the kernel layer (two launches per layer for all of them) + one async count read-back.
  void NextGroup(std::vector<TaskPtr> *group);
  void EnqueueGroup(const std::vector<TaskPtr> &group);
  void Enqueue(const TaskPtr &task) { EnqueueGroup({task}); }
  void Finish(const TaskPtr &task);               // the ONE host sync of the batch + exact-size tensors
  void Sample(const TaskPtr &task) { Enqueue(task); Finish(task); }
  size_t GroupSize() const { return group_; }
  void CountFrequency(const TaskPtr &task, uint32_t *d_freq);  // PreSC: freq[input_nodes] += 1 (after Enqueue)
  void SyncAll();
  bool Done(const TaskPtr &task) { return cudaEventQuery(slots_[task->slot].done) == cudaSuccess; }
  void SyncSlot(const TaskPtr &task) { CUDA_CALL(cudaStreamSynchronize(slots_[task->slot].stream)); }
  void ResetShuffler() {
    cur_epoch_ = 0; cur_step_ = 0; shuffled_epoch_ = (uint64_t)-1;
    next_slot_ = (next_slot_ + group_ - 1) / group_ * group_;  // slot groups stay aligned
  }
  // PreSC draws its neighbours from its own Philox stream.  The reference pre-samples with the live cuRAND states
  // and then rewinds only the shuffler (pre_sampler.cc:101-103), so its training epochs never repeat the
  // pre-sampling draws; with a counter-based RNG keyed by (seed, batch key) they would be repeated exactly and
  // epoch 0 would see an optimistic cache hit rate (every node it touches was counted).
  void SetRngSalt(uint64_t salt) { rng_salt_ = salt; }
  size_t NumStep() const { return num_step_; }
  size_t NumLocalStep() const { return local_steps_; }
  size_t NumSlots() const { return slots_.size(); }
  cudaStream_t stream() const { return stream_; }
  int device() const { return dev_; }
  const IdType *d_indptr() const { return (const IdType *)indptr_->data; }
  const IdType *d_indices() const { return (const IdType *)indices_->data; }
  size_t max_nodes() const { return max_nodes_; }

 private:
  void Reshuffle(uint64_t epoch);
  const Dataset *ds_;
  int dev_;
  cudaStream_t stream_ = nullptr;  // uploads + epoch shuffle
  cudaEvent_t shuffled_ = nullptr;
  RunConfig &rc_;
  std::vector<size_t> fanout_;
  size_t L_, batch_;
  // topology + weights in HBM (dist_engine.cc:176-191)
  TensorPtr indptr_, indices_, prob_, alias_, prefix_;
  // shuffler (dist_shuffler.cc:37-96 split; permutation drawn on the GPU, shuffle.cu)
  TensorPtr train_dev_, perm_dev_, shuffle_ws_;
  size_t num_train_, num_step_, local_steps_, step_begin_;
  uint64_t cur_epoch_ = 0, cur_step_ = 0, shuffled_epoch_ = (uint64_t)-1, num_epoch_;
  uint64_t rng_salt_ = 0;
  std::vector<uint8_t> sanity_map_;  // SAMGRAPH_SANITY_CHECK: train nodes already handed out this epoch
  // hash table + scratch sized from PredictNumNodes, one set per slot
  size_t max_nodes_, ht_cap_;
  std::vector<size_t> in_max_, edge_max_;
  std::vector<SampleSlot> slots_;
  size_t next_slot_ = 0, group_ = 1;
  TensorPtr counts_dev_, counts_host_;  // [slot][3L+1], contiguous: one read-back per group
  std::shared_ptr<BlockPool> pool_;
};

Sampler::Sampler(const Dataset *ds, Context ctx, int worker_id, int num_worker, size_t num_epoch)
    : ds_(ds), dev_(ctx.device_id), rc_(RunConfig::Get()), num_epoch_(num_epoch) {
  FCHECK(ctx.device_type == kGPU) << "the sampler must live on a GPU: there is no CPU sampling path";
  CUDA_CALL(cudaSetDevice(dev_));
  CUDA_CALL(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  CUDA_CALL(cudaEventCreateWithFlags(&shuffled_, cudaEventDisableTiming));
  fanout_ = rc_.fanout;
  L_ = fanout_.size();
  batch_ = rc_.batch_size;

  auto upload = [&](const TensorPtr &h, const char *name) -> TensorPtr {
    if (!h || !h->data) return nullptr;
    auto d = Tensor::Device(h->dtype, h->shape, dev_, stream_, name);
    CUDA_CALL(cudaMemcpyAsync(d->data, h->data, h->nbytes, cudaMemcpyHostToDevice, stream_));
    return d;
  };
  indptr_ = upload(ds->indptr, "indptr");
  indices_ = upload(ds->indices, "indices");
  prob_ = upload(ds->prob_table, "prob_table");
  alias_ = upload(ds->alias_table, "alias_table");
  prefix_ = upload(ds->prob_prefix_table, "prob_prefix_table");

  // ---- shuffler split (dist_shuffler.cc:60-83): worker w owns steps [w*floor(N/S), ...) ----
  num_train_ = ds->train_set->NumItems();
  num_step_ = (num_train_ + batch_ - 1) / batch_;
  fgnn_rt_step_split(num_step_, (size_t)num_worker, (size_t)worker_id, &step_begin_, &local_steps_);
  train_dev_ = upload(ds->train_set, "train_set");
  perm_dev_ = Tensor::Device(kI32, {num_train_}, dev_, stream_, "train_perm");
  shuffle_ws_ = Tensor::Device(kU8, {fgnn_k_shuffle_workspace_bytes(num_train_)}, dev_, stream_, "shuffle_ws");

  // ---- per-batch scratch at the PredictNumNodes bounds (common.cc:330-339) ----
  max_nodes_ = PredictNumNodes(batch_, fanout_, L_);
  ht_cap_ = fgnn_k_ht_capacity(max_nodes_);
  in_max_.resize(L_);
  edge_max_.resize(L_);
  size_t cur = batch_;
  size_t ws_bytes = 16;
  for (int i = (int)L_ - 1; i >= 0; --i) {
    in_max_[i] = cur;
    edge_max_[i] = cur * fanout_[i];
    FCHECK_LT(edge_max_[i], (size_t)0x7FFFFFFF) << "layer too large for 32-bit edge counts";
    cur += cur * fanout_[i];
    if (rc_.sample_type == kKHop1 || rc_.sample_type == kWeightedKHop || rc_.sample_type == kWeightedKHopPrefix)
      ws_bytes = std::max(ws_bytes, fgnn_k_sample_replace_workspace_bytes((uint32_t)in_max_[i], (uint32_t)fanout_[i]));
    if (rc_.sample_type == kRandomWalk)
      ws_bytes = std::max(ws_bytes, fgnn_k_sample_random_walk_workspace_bytes((uint32_t)in_max_[i], (uint32_t)fanout_[i]));
  }
  // Super-batch (round 2): `group_` mini-batches are enqueued together, every layer being two launches for all of
  // them (fgnn_k_sample_batch_multi), on ONE stream per slot group; two groups alternate so that one is sampled
  // while the other is consumed.  FGNN_SUPER_BATCH=1 restores round 1's one-batch-per-stream slots.
  group_ = 4;
  if (!GetEnv("FGNN_SUPER_BATCH").empty()) group_ = (size_t)std::max(1, atoi(GetEnv("FGNN_SUPER_BATCH").c_str()));
  group_ = std::min<size_t>(group_, FGNN_MAX_SUPER);
  size_t nslots = group_ > 1 ? 2 * group_ : 4;
  // numeric / "0" knobs are read by value: IsEnvSet() is only true for 1/ON/On/on (common.cc semantics)
  if (group_ == 1 && !GetEnv("FGNN_SAMPLER_SLOTS").empty())
    nslots = (size_t)std::min(8, std::max(1, atoi(GetEnv("FGNN_SAMPLER_SLOTS").c_str())));
  slots_.resize(nslots);
  const size_t cw = L_ * 3 + 1;
  counts_dev_ = Tensor::Device(kI32, {nslots * cw}, dev_, stream_, "counts");
  counts_host_ = Tensor::Pinned(kI32, {nslots * cw}, "counts_host");
  for (size_t si = 0; si < nslots; ++si) {
    SampleSlot &sl = slots_[si];
    if (si % group_ == 0) CUDA_CALL(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
    else sl.stream = slots_[si - si % group_].stream;  // a slot group shares one stream
    CUDA_CALL(cudaEventCreateWithFlags(&sl.done, TraceGpu() ? cudaEventDefault : cudaEventDisableTiming));
    CUDA_CALL(cudaEventCreateWithFlags(&sl.idle, cudaEventDisableTiming));
    if (TraceGpu()) CUDA_CALL(cudaEventCreate(&sl.begin));
    sl.table = Tensor::Device(kU8, {fgnn_k_ht_bytes(ht_cap_)}, dev_, sl.stream, "hashtable");
    sl.num_items = Tensor::Device(kI32, {4}, dev_, sl.stream, "num_items");
    sl.chain = Tensor::Device(kU8, {FGNN_CHAIN_WS_BYTES}, dev_, sl.stream, "chain_ws");
    CUDA_CALL(cudaMemsetAsync(sl.chain->data, 0, FGNN_CHAIN_WS_BYTES, sl.stream));
    sl.counts_dev = (uint32_t *)counts_dev_->data + si * cw;
    sl.counts_host = (uint32_t *)counts_host_->data + si * cw;
    sl.ws = Tensor::Device(kU8, {ws_bytes}, dev_, sl.stream, "sample_ws");
    if ((rc_.sample_type == kKHop1 || rc_.sample_type == kWeightedKHop || rc_.sample_type == kWeightedKHopPrefix) &&
        GetEnv("FGNN_SEED_RANK") != "0") {
      // seeds ordered by id with a rank-by-bitmap (sample_replace.cu) instead of a library radix sort
      const size_t rb = fgnn_k_seed_rank_workspace_bytes(ds->num_node);
      if (rb) {
        sl.rank_ws = Tensor::Device(kU8, {rb}, dev_, sl.stream, "seed_rank_ws");
        CUDA_CALL(cudaMemsetAsync(sl.rank_ws->data, 0, rb, sl.stream));
      }
    }
    for (size_t i = 0; i < L_; ++i) {
      sl.dst.push_back(Tensor::Device(kI32, {edge_max_[i] + 1}, dev_, sl.stream, "scratch_dst"));
      sl.pos.push_back(Tensor::Device(kI32, {edge_max_[i] + 1}, dev_, sl.stream, "scratch_pos"));
    }
    CUDA_CALL(cudaStreamSynchronize(sl.stream));
  }
  pool_ = std::make_shared<BlockPool>(dev_, batch_, max_nodes_, edge_max_, rc_.sample_type == kRandomWalk);
  pool_->Reserve(nslots + rc_.max_sampling_jobs + rc_.max_copying_jobs + 2);
  if (rc_.sample_type == kWeightedKHop || rc_.sample_type == kWeightedKHopHashDedup)
    FCHECK(prob_ && alias_) << "weighted sampling needs prob_table.bin and alias_table.bin";
  if (rc_.sample_type == kWeightedKHopPrefix) FCHECK(prefix_) << "needs prob_prefix_table.bin";
  CUDA_CALL(cudaStreamSynchronize(stream_));
}

void Sampler::SyncAll() {
  cudaSetDevice(dev_);
  for (auto &sl : slots_)
    if (sl.stream) cudaStreamSynchronize(sl.stream);
  if (stream_) cudaStreamSynchronize(stream_);
}

Sampler::~Sampler() {
  SyncAll();  // tensors free themselves on their streams afterwards
  cudaGetLastError();
}

void Sampler::Reshuffle(uint64_t epoch) {
  SlowScope ss("Sampler::Reshuffle");
  // batches still in flight read the old permutation: order the shuffle after them
  for (auto &sl : slots_) {
    CUDA_CALL(cudaEventRecord(sl.idle, sl.stream));
    CUDA_CALL(cudaStreamWaitEvent(stream_, sl.idle, 0));
  }
  FGNN_CALL(fgnn_k_shuffle((const IdType *)train_dev_->data, num_train_, rc_.seed, epoch,
                           (IdType *)perm_dev_->data, shuffle_ws_->data, shuffle_ws_->nbytes,
                           (fgnn_stream_t)stream_));
  CUDA_CALL(cudaEventRecord(shuffled_, stream_));
  for (auto &sl : slots_) CUDA_CALL(cudaStreamWaitEvent(sl.stream, shuffled_, 0));
  shuffled_epoch_ = epoch;
  if (rc_.option_sanity_check) sanity_map_.assign(ds_->num_node, 0);  // one epoch = every train node at most once
}

TaskPtr Sampler::Next() {
  if (cur_step_ >= local_steps_) {
    cur_step_ = 0;
    ++cur_epoch_;
  }
  if (cur_epoch_ >= num_epoch_) return nullptr;
  CUDA_CALL(cudaSetDevice(dev_));
  if (shuffled_epoch_ != cur_epoch_) Reshuffle(cur_epoch_);
  const size_t gstep = step_begin_ + cur_step_;
  const size_t off = gstep * batch_;
  const size_t n = std::min(batch_, num_train_ - off);
  SlowScope ss("Sampler::Next");
  auto task = std::make_shared<Task>();
  task->key = cur_epoch_ * num_step_ + gstep;  // global key (dist_loops.cc:41-43)
  task->slot = (int)(next_slot_++ % slots_.size());
  cudaStream_t st = slots_[task->slot].stream;
  // Copy1D slice of the permuted train set (cuda_shuffler.cc:128-154) into the batch's block
  task->block = pool_->Acquire();
  TaskBlock *blk = static_cast<TaskBlock *>(task->block.get());
  task->output_nodes = Tensor::View(blk->seeds, kI32, {n}, Context(kGPU, dev_), task->block, "output_nodes");
  CUDA_CALL(cudaMemcpyAsync(blk->seeds, (const IdType *)perm_dev_->data + off, n * sizeof(IdType),
                            cudaMemcpyDeviceToDevice, st));
  if (rc_.option_sanity_check) {  // cuda_shuffler.cc:144-151 (debug switch: one small D2H copy + host scan per batch)
    std::vector<IdType> h(n);
    CUDA_CALL(cudaMemcpyAsync(h.data(), blk->seeds, n * sizeof(IdType), cudaMemcpyDeviceToHost, st));
    CUDA_CALL(cudaStreamSynchronize(st));
    size_t bad = 0;
    const int rc = fgnn_rt_sanity_check_batch(sanity_map_.data(), ds_->num_node, h.data(), n, &bad);
    FCHECK(rc == 0) << (rc == 3 ? "duplicate batch input" : rc == 1 ? "empty key in batch input" : "batch input out of range")
                    << ": seed " << (bad < n ? h[bad] : 0) << " at position " << bad << " of batch " << task->key;
  }
  ++cur_step_;
  return task;
}

void Sampler::NextGroup(std::vector<TaskPtr> *group) {
  group->clear();
  next_slot_ = (next_slot_ + group_ - 1) / group_ * group_;  // a group never straddles two slot groups
  while (group->size() < group_) {
    TaskPtr t = Next();
    if (!t) break;
    group->push_back(t);
  }
}

void Sampler::EnqueueGroup(const std::vector<TaskPtr> &group) {
  SlowScope ss("Sampler::EnqueueGroup");
  if (group.empty()) return;
  CUDA_CALL(cudaSetDevice(dev_));
  const size_t K = group.size();
  FCHECK_LE(K, (size_t)FGNN_MAX_SUPER);
  const size_t first_slot = (size_t)group[0]->slot;
  cudaStream_t stream = slots_[first_slot].stream;
  fgnn_stream_t st = (fgnn_stream_t)stream;
  const bool versioned = (rc_.sample_type == kKHop2) && GetEnv("FGNN_HT_VERSIONED") != "0";
  if (rc_.sample_type == kRandomWalk)
    for (size_t i = 0; i < L_; ++i) FCHECK_EQ(fanout_[i], rc_.num_neighbor);

  // the whole of DoGPUSample (cuda_loops.cc:50-267) for every mini-batch of the group is one call into the
  // kernel layer: per layer sample(+insert) -> compact(+remap); nothing below waits on the host
  fgnn_sample_plan pl[FGNN_MAX_SUPER];
  fgnn_sample_out so[FGNN_MAX_SUPER];
  const fgnn_sample_plan *plp[FGNN_MAX_SUPER];
  const fgnn_sample_out *sop[FGNN_MAX_SUPER];
  const uint32_t *seeds[FGNN_MAX_SUPER];
  uint32_t n_seed[FGNN_MAX_SUPER];
  uint64_t keys[FGNN_MAX_SUPER];
  memset(pl, 0, sizeof(pl));
  memset(so, 0, sizeof(so));
  for (size_t k = 0; k < K; ++k) {
    const TaskPtr &task = group[k];
    FCHECK(task->slot >= 0 && (size_t)task->slot == first_slot + k && first_slot / group_ == (first_slot + k) / group_)
        << "a super-batch uses consecutive slots of one slot group";
    SampleSlot &sl = slots_[task->slot];
    TaskBlock *blk = static_cast<TaskBlock *>(task->block.get());
    if (sl.begin) CUDA_CALL(cudaEventRecord(sl.begin, stream));
    fgnn_sample_plan &p = pl[k];
    p.sample_type = (int32_t)rc_.sample_type;
    p.num_layers = (uint32_t)L_;
    for (size_t i = 0; i < L_; ++i) {
      p.fanout[i] = (uint32_t)fanout_[i];
      p.in_max[i] = (uint32_t)in_max_[i];
      p.dst[i] = (uint32_t *)sl.dst[i]->data;
      p.pos[i] = (uint32_t *)sl.pos[i]->data;
    }
    p.indptr = d_indptr();
    p.indices = d_indices();
    p.prob_table = prob_ ? (const float *)prob_->data : nullptr;
    p.alias_table = alias_ ? (const IdType *)alias_->data : nullptr;
    p.prob_prefix_table = prefix_ ? (const float *)prefix_->data : nullptr;
    p.walk_len = (uint32_t)rc_.random_walk_length;
    p.num_walk = (uint32_t)rc_.num_random_walk;
    p.restart_prob = rc_.random_walk_restart_prob;
    p.seed = rc_.seed ^ rng_salt_;
    p.table = sl.table->data;
    p.capacity = ht_cap_;
    p.num_items = (uint32_t *)sl.num_items->data;
    p.chain_ws = sl.chain->data;
    p.workspace = sl.ws->data;
    p.workspace_bytes = sl.ws->nbytes;
    p.rank_ws = sl.rank_ws ? sl.rank_ws->data : nullptr;
    p.num_nodes = sl.rank_ws ? (uint32_t)ds_->num_node : 0u;
    p.version = versioned ? fgnn_k_ht_next_version(&sl.ver_state, p.table, ht_cap_, st) : 0u;
    fgnn_sample_out &o = so[k];
    o.n2o = blk->n2o;
    o.counts = sl.counts_dev;  // [L][3] = num_dst, num_edge, num_src
    for (size_t i = 0; i < L_; ++i) {
      o.row[i] = blk->row[i];
      o.col[i] = blk->col[i];
      o.data[i] = blk->data[i];
    }
    plp[k] = &p;
    sop[k] = &o;
    seeds[k] = (const IdType *)task->output_nodes->data;
    n_seed[k] = (uint32_t)task->output_nodes->NumItems();
    keys[k] = task->key;
  }
  FGNN_CALL(fgnn_k_sample_batch_multi(plp, sop, seeds, n_seed, keys, (uint32_t)K, st));
  // all counts of the group go to the host at once (the slots' count arrays are contiguous)
  const size_t cw = L_ * 3 + 1;
  CUDA_CALL(cudaMemcpyAsync(slots_[first_slot].counts_host, slots_[first_slot].counts_dev, K * cw * 4,
                            cudaMemcpyDeviceToHost, stream));
  for (size_t k = 0; k < K; ++k) CUDA_CALL(cudaEventRecord(slots_[first_slot + k].done, stream));
}

void Sampler::Finish(const TaskPtr &task) {
  SlowScope ss("Sampler::Finish");
  CUDA_CALL(cudaSetDevice(dev_));
  SampleSlot &sl = slots_[task->slot];
  cudaStream_t stream = sl.stream;
  CUDA_CALL(cudaEventSynchronize(sl.done));  // the ONE host round trip of the batch
  if (sl.begin) {
    float ms = 0.f;
    CUDA_CALL(cudaEventElapsedTime(&ms, sl.begin, sl.done));
    Profiler::Get().LogStep(task->key, kLogL2IdCopyTime, ms * 1e-3);  // trace: device time of the sampling chain
  }
  const uint32_t *h = sl.counts_host;

  // exact-size views of the batch's block (TrainGraph: row = neighbour local id, col = seed local id, :210-229)
  TaskBlock *blk = static_cast<TaskBlock *>(task->block.get());
  const Context ctx(kGPU, dev_);
  task->graphs.resize(L_);
  size_t total_edges = 0;
  for (size_t i = 0; i < L_; ++i) {
    TrainGraph &g = task->graphs[i];
    g.num_dst = h[3 * i];
    g.num_edge = h[3 * i + 1];
    g.num_src = h[3 * i + 2];
    total_edges += g.num_edge;
    g.row = Tensor::View(blk->row[i], kI32, {g.num_edge}, ctx, task->block, "train_graph.row");
    g.col = Tensor::View(blk->col[i], kI32, {g.num_edge}, ctx, task->block, "train_graph.col");
    if (blk->data[i]) g.data = Tensor::View(blk->data[i], kI32, {g.num_edge}, ctx, task->block, "train_graph.data");
  }
  const size_t n_input = h[2];  // num_src of layer 0 == number of unique nodes
  task->input_nodes = Tensor::View(blk->n2o, kI32, {n_input}, ctx, task->block, "input_nodes");
  if (!task->ready) CUDA_CALL(cudaEventCreateWithFlags(&task->ready, cudaEventDisableTiming));
  CUDA_CALL(cudaEventRecord(task->ready, stream));
  Profiler::Get().LogStep(task->key, kLogL1NumNode, (double)n_input);
  Profiler::Get().LogStep(task->key, kLogL1NumSample, (double)total_edges);
}

void Sampler::CountFrequency(const TaskPtr &task, uint32_t *d_freq) {
  SampleSlot &sl = slots_[task->slot];
  FGNN_CALL(fgnn_k_freq_count(d_freq, static_cast<TaskBlock *>(task->block.get())->n2o, (uint32_t)max_nodes_,
                              (const uint32_t *)sl.num_items->data, (fgnn_stream_t)sl.stream));
}

// =============================================================================================
// Extractor: DoGraphCopy + DoCacheFeatureCopy + label extract on the trainer GPU
// (cuda_loops.cc:599-606, dist_loops.cc:713-929, dist_cache_manager_*.{cc,cu})
// =============================================================================================
class Extractor {
 public:
  Extractor(const Dataset *ds, Context ctx, const IdType *ranking_host_or_null, const IdType *ranking_dev_or_null,
            int ranking_dev, int shard_id, int num_shards, SharedRing *ring);
  ~Extractor();
  // task tensors must already live on this device
  void Enqueue(const TaskPtr &task);   // gather kernels + async stats read-back, no host sync
  void Finish(const TaskPtr &task);    // wait for the batch, hit/miss accounting
  void Extract(const TaskPtr &task) { Enqueue(task); Finish(task); }
  bool Done(const TaskPtr &task) { return cudaEventQuery(task->extracted) == cudaSuccess; }
  static constexpr int kDepth = 2;     // batches in flight on the extraction stream
  TaskPtr MoveToTrainer(const TaskPtr &task, int src_dev);
  cudaStream_t stream() const { return stream_; }
  int device() const { return dev_; }

 private:
  const Dataset *ds_;
  int dev_;
  cudaStream_t stream_ = nullptr;
  RunConfig &rc_;
  size_t row_bytes_, num_cached_ = 0;
  uint64_t feat_mask_ = ~0ull;
  const void *feat_src_ = nullptr;     // pinned / registered host feature table (UVA)
  bool feat_registered_ = false;
  TensorPtr feat_pinned_, label_dev_, cache_table_, shard_ptrs_, stats_, stats_host_, defer_ws_;
  unsigned long long last_stats_[2] = {0, 0};
  uint64_t enq_seq_ = 0;
  void *shard_ = nullptr;              // this GPU's stripe of the cache rows (cudaMalloc: IPC exportable)
  void *replica_ = nullptr;            // hybrid layout: the hottest rows, held by every trainer
  size_t num_replicated_ = 0;
  fgnn_cache_layout layout_;
  std::vector<void *> peer_shards_;
  int num_shards_ = 1, shard_id_ = 0;
};

Extractor::Extractor(const Dataset *ds, Context ctx, const IdType *ranking_host, const IdType *ranking_dev,
                     int ranking_dev_id, int shard_id, int num_shards, SharedRing *ring)
    : ds_(ds), dev_(ctx.device_id), rc_(RunConfig::Get()), num_shards_(num_shards), shard_id_(shard_id) {
  FCHECK(ctx.device_type == kGPU) << "the trainer must be a GPU";
  CUDA_CALL(cudaSetDevice(dev_));
  {
    // the HBM-bound gather runs at the highest stream priority: its CTAs (one 128 KB shared-memory ring per
    // SM) must not queue behind the many small CTAs of the latency-bound sampling kernels of other slots
    int lo = 0, hi = 0;
    CUDA_CALL(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const bool prio = GetEnv("FGNN_EXTRACT_PRIORITY") != "0";
    CUDA_CALL(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, prio ? hi : lo));
  }
  const size_t V = ds->num_node, D = ds->feat_dim;
  row_bytes_ = D * DataTypeBytes(ds->feat->dtype);

  // ---- host feature table readable from the GPU (miss path: UVA loads over the host link) ----
  if (rc_.option_empty_feat != 0) feat_mask_ = (1ull << rc_.option_empty_feat) - 1;  // cuda_extraction.cu:131
  if (ds->feat->ctx.device_type == kMMAP) {
    cudaError_t e = cudaHostRegister(ds->feat->data, ds->feat->nbytes,
                                     cudaHostRegisterMapped | cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    if (e == cudaSuccess) {
      feat_registered_ = true;
      feat_src_ = ds->feat->data;
    } else {
      cudaGetLastError();
      FLOG(Warning) << "cudaHostRegister(feat.bin) failed (" << cudaGetErrorString(e)
                    << "); staging the feature table into pinned memory";
      feat_pinned_ = Tensor::Pinned(ds->feat->dtype, ds->feat->shape, "feat_pinned");
      memcpy(feat_pinned_->data, ds->feat->data, ds->feat->nbytes);
      feat_src_ = feat_pinned_->data;
    }
  } else {
    feat_src_ = ds->feat->data;  // already pinned (EmptyNoScale path, engine.cc:139-155)
  }

  // ---- labels live in HBM; the reference gathers them on the CPU (dist_loops.cc:886-929) ----
  label_dev_ = Tensor::Device(kI64, {V}, dev_, stream_, "label");
  CUDA_CALL(cudaMemcpyAsync(label_dev_->data, ds->label->data, V * 8, cudaMemcpyHostToDevice, stream_));

  stats_ = Tensor::Device(kI64, {4}, dev_, stream_, "gather_stats");  // hits, misses, rows read from peer shards
  CUDA_CALL(cudaMemsetAsync(stats_->data, 0, 32, stream_));
  stats_host_ = Tensor::Pinned(kI64, {2 * kDepth}, "gather_stats_host");

  // ---- cache: node -> slot table + the rows of this shard -------------------------------------
  const bool full_gpu = (rc_.run_arch == kArch1);  // arch1: every feature row is HBM resident
  const double pct = full_gpu ? 1.0 : rc_.cache_percentage;
  num_cached_ = (size_t)((double)V * pct);         // dist_cache_manager_host.cc:66
  if (num_cached_ > V) num_cached_ = V;
  cache_table_ = Tensor::Device(kI32, {V}, dev_, stream_, "cache_table");
  TensorPtr rank_dev;
  const IdType *rank = nullptr;
  if (full_gpu) {
    // identity ranking: slot == node id (no PreSC needed)
    rank_dev = Tensor::Device(kI32, {V}, dev_, stream_, "rank_identity");
    std::vector<IdType> ident(V);
    for (size_t i = 0; i < V; ++i) ident[i] = (IdType)i;
    CUDA_CALL(cudaMemcpyAsync(rank_dev->data, ident.data(), V * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_CALL(cudaStreamSynchronize(stream_));
    rank = (const IdType *)rank_dev->data;
  } else if (num_cached_ > 0) {
    rank_dev = Tensor::Device(kI32, {V}, dev_, stream_, "ranking_nodes");
    if (ranking_dev && ranking_dev_id == dev_) {
      CUDA_CALL(cudaMemcpyAsync(rank_dev->data, ranking_dev, V * 4, cudaMemcpyDeviceToDevice, stream_));
    } else if (ranking_dev) {
      CUDA_CALL(cudaMemcpyPeerAsync(rank_dev->data, dev_, ranking_dev, ranking_dev_id, V * 4, stream_));
    } else {
      FCHECK(ranking_host) << "cache enabled but no ranking available";
      CUDA_CALL(cudaMemcpyAsync(rank_dev->data, ranking_host, V * 4, cudaMemcpyHostToDevice, stream_));
    }
    rank = (const IdType *)rank_dev->data;
  }
  FGNN_CALL(fgnn_k_cache_table_build((uint32_t *)cache_table_->data, V, rank, num_cached_, (fgnn_stream_t)stream_));

  // Hybrid layout (num_shards > 1): the hottest R = replicate_percentage * V ranks are kept by EVERY trainer
  // (with a power-law hotness ranking they take most of the hits off the NVLink path); the slots behind them are
  // striped: slot s >= R lives on trainer (s-R) % T at local row (s-R) / T.  Each trainer fills only its own
  // replica and stripe from the host table (the reference has every trainer gather ALL cached rows,
  // dist_cache_manager_host.cc:98-109).
  if (num_shards > 1) {
    num_replicated_ = std::min(num_cached_, (size_t)((double)V * std::max(0.0, rc_.replicate_percentage)));
    if (num_replicated_) {
      FGNN_CALL(fgnn_k_shard_alloc(&replica_, num_replicated_ * row_bytes_));
      FGNN_CALL(fgnn_k_row_copy(replica_, nullptr, feat_src_, rank, feat_mask_, (uint32_t)num_replicated_, nullptr,
                                row_bytes_, (fgnn_stream_t)stream_));
    }
  }
  const size_t striped = num_cached_ - num_replicated_;
  const size_t local_rows = striped > (size_t)shard_id ? (striped - shard_id + num_shards - 1) / num_shards : 0;
  FGNN_CALL(fgnn_k_shard_alloc(&shard_, std::max<size_t>(local_rows, 1) * row_bytes_));
  if (local_rows) {
    if (num_shards == 1) {
      FGNN_CALL(fgnn_k_row_copy(shard_, nullptr, feat_src_, rank, feat_mask_, (uint32_t)local_rows, nullptr, row_bytes_,
                                (fgnn_stream_t)stream_));
    } else {
      // strided slice of the ranking: gather ids first
      auto ids = Tensor::Device(kI32, {local_rows}, dev_, stream_, "shard_ids");
      CUDA_CALL(cudaMemcpy2DAsync(ids->data, 4, rank + num_replicated_ + shard_id, (size_t)num_shards * 4, 4, local_rows,
                                  cudaMemcpyDeviceToDevice, stream_));
      FGNN_CALL(fgnn_k_row_copy(shard_, nullptr, feat_src_, (const IdType *)ids->data, feat_mask_, (uint32_t)local_rows,
                                nullptr, row_bytes_, (fgnn_stream_t)stream_));
      CUDA_CALL(cudaStreamSynchronize(stream_));
    }
  }
  CUDA_CALL(cudaStreamSynchronize(stream_));

  // ---- map the peers' shards (CUDA IPC across trainer processes, loads go over NVLink) ----
  peer_shards_.assign(num_shards, nullptr);
  peer_shards_[shard_id] = shard_;
  if (num_shards > 1) {
    FCHECK(ring) << "partitioned cache needs the shared segment";
    FCHECK_LE(num_shards, 16);
    FGNN_CALL(fgnn_k_ipc_export(shard_, ring->ipc_handle[shard_id]));
    ring->shard_rows[shard_id] = local_rows;
    pthread_barrier_wait(&ring->trainer_barrier);
    for (int t = 0; t < num_shards; ++t)
      if (t != shard_id) FGNN_CALL(fgnn_k_ipc_open(ring->ipc_handle[t], &peer_shards_[t]));
    pthread_barrier_wait(&ring->trainer_barrier);
  }
  shard_ptrs_ = Tensor::Device(kI64, {(size_t)num_shards}, dev_, stream_, "shard_ptrs");
  CUDA_CALL(cudaMemcpyAsync(shard_ptrs_->data, peer_shards_.data(), num_shards * sizeof(void *), cudaMemcpyHostToDevice,
                            stream_));
  CUDA_CALL(cudaStreamSynchronize(stream_));
  memset(&layout_, 0, sizeof(layout_));
  layout_.table = (const uint32_t *)cache_table_->data;
  layout_.shards = (const void *const *)shard_ptrs_->data;
  layout_.num_shards = (uint32_t)num_shards;
  layout_.self_shard = (uint32_t)shard_id;
  layout_.replica = replica_;
  layout_.num_replicated = (uint32_t)num_replicated_;
  layout_.miss_src = feat_src_;
  layout_.miss_mask = feat_mask_;
  layout_.row_bytes = row_bytes_;
  if (num_shards > 1) {  // peer rows in a second pass (fgnn_cache_layout.defer_ws)
    const size_t bound = PredictNumNodes(rc_.batch_size, rc_.fanout, rc_.fanout.size());
    const size_t nb = fgnn_k_gather_defer_workspace_bytes((uint32_t)std::min<size_t>(bound, 0xFFFFFFFFu));
    defer_ws_ = Tensor::Device(kU8, {nb}, dev_, stream_, "gather_defer_ws");
    CUDA_CALL(cudaMemsetAsync(defer_ws_->data, 0, 256, stream_));
    CUDA_CALL(cudaStreamSynchronize(stream_));
    layout_.defer_ws = defer_ws_->data;
  }
  FLOG(Info) << "GPU cache: " << num_cached_ << " / " << V << " nodes, " << num_replicated_ << " replicated, shard "
             << shard_id << "/" << num_shards << " holds " << local_rows << " rows ("
             << (local_rows * row_bytes_ >> 20) << " MiB)";
}

Extractor::~Extractor() {
  cudaSetDevice(dev_);
  if (stream_) cudaStreamSynchronize(stream_);
  for (int t = 0; t < (int)peer_shards_.size(); ++t)
    if (t != shard_id_ && peer_shards_[t]) fgnn_k_ipc_close(peer_shards_[t]);
  if (shard_) fgnn_k_shard_free(shard_);
  if (replica_) fgnn_k_shard_free(replica_);
  if (feat_registered_) cudaHostUnregister(const_cast<void *>(feat_src_));
  cudaGetLastError();
}

TaskPtr Extractor::MoveToTrainer(const TaskPtr &task, int src_dev) {  // DoGraphCopy, cuda_loops.cc:599-606
  if (src_dev == dev_) return task;
  CUDA_CALL(cudaSetDevice(dev_));
  if (task->ready) CUDA_CALL(cudaStreamWaitEvent(stream_, task->ready, 0));
  auto move = [&](TensorPtr &t) {
    if (!t) return;
    auto d = Tensor::Device(t->dtype, t->shape, dev_, stream_, t->name);
    CUDA_CALL(cudaMemcpyPeerAsync(d->data, dev_, t->data, src_dev, t->nbytes, stream_));
    t = d;
  };
  auto out = std::make_shared<Task>();
  out->key = task->key;
  out->graphs = task->graphs;
  out->input_nodes = task->input_nodes;
  out->output_nodes = task->output_nodes;
  for (auto &g : out->graphs) { move(g.row); move(g.col); move(g.data); }
  move(out->input_nodes);
  move(out->output_nodes);
  CUDA_CALL(cudaStreamSynchronize(stream_));  // sources may be released after this
  return out;
}

void Extractor::Enqueue(const TaskPtr &task) {
  SlowScope ss("Extractor::Enqueue");
  CUDA_CALL(cudaSetDevice(dev_));
  if (task->ready) CUDA_CALL(cudaStreamWaitEvent(stream_, task->ready, 0));
  const size_t n_in = task->input_nodes->NumItems(), n_out = task->output_nodes->NumItems();
  const size_t D = ds_->feat_dim;
  {
    SlowScope sa("Extractor::Enqueue alloc");
    task->input_feat = Tensor::Device(ds_->feat->dtype, {n_in, D}, dev_, stream_, "input_feat");
    task->output_label = Tensor::Device(kI64, {n_out}, dev_, stream_, "output_label");
  }
  task->xslot = (int)(enq_seq_++ % kDepth);
  unsigned long long *after = (unsigned long long *)stats_host_->data + 2 * task->xslot;
  if (TraceGpu()) {
    if (!task->xbegin) CUDA_CALL(cudaEventCreate(&task->xbegin));
    CUDA_CALL(cudaEventRecord(task->xbegin, stream_));
  }
  // one fused kernel instead of GetMissCacheIndex + ExtractMissData + H2D + 2 combine kernels
  FGNN_CALL(fgnn_k_gather_cached_layout(task->input_feat->data, (const IdType *)task->input_nodes->data, (uint32_t)n_in,
                                        nullptr, &layout_, (unsigned long long *)stats_->data,
                                        (unsigned long long *)stats_->data + 2, (fgnn_stream_t)stream_));
  // labels: GPUExtract with D = 1, int64 (cuda_extraction.cu:74-117)
  FGNN_CALL(fgnn_k_row_copy(task->output_label->data, nullptr, label_dev_->data, (const IdType *)task->output_nodes->data,
                            ~0ull, (uint32_t)n_out, nullptr, 8, (fgnn_stream_t)stream_));
  CUDA_CALL(cudaMemcpyAsync(after, stats_->data, 16, cudaMemcpyDeviceToHost, stream_));
  if (!task->extracted)
    CUDA_CALL(cudaEventCreateWithFlags(&task->extracted, TraceGpu() ? cudaEventDefault : cudaEventDisableTiming));
  CUDA_CALL(cudaEventRecord(task->extracted, stream_));
}

void Extractor::Finish(const TaskPtr &task) {
  SlowScope ss("Extractor::Finish");
  CUDA_CALL(cudaSetDevice(dev_));
  CUDA_CALL(cudaEventSynchronize(task->extracted));
  if (task->xbegin) {
    float ms = 0.f;
    CUDA_CALL(cudaEventElapsedTime(&ms, task->xbegin, task->extracted));
    Profiler::Get().LogStep(task->key, kLogL2ExtractTime, ms * 1e-3);  // trace: device time of gather + labels
  }
  const unsigned long long *after = (const unsigned long long *)stats_host_->data + 2 * task->xslot;
  const size_t n_in = task->input_nodes->NumItems(), n_out = task->output_nodes->NumItems();
  task->num_cache = after[0] - last_stats_[0];   // batches finish in enqueue order: cumulative counters
  task->num_miss = after[1] - last_stats_[1];
  last_stats_[0] = after[0];
  last_stats_[1] = after[1];
  FCHECK_EQ(task->num_miss + task->num_cache, n_in);  // cuda_loops.cc:999 invariant
  auto &p = Profiler::Get();
  p.LogStep(task->key, kLogL1FeatureBytes, (double)(n_in * row_bytes_));
  p.LogStep(task->key, kLogL1LabelBytes, (double)(n_out * 8));
  p.LogStep(task->key, kLogL1MissBytes, (double)(task->num_miss * row_bytes_));
  p.LogEpochAdd(task->key, kLogEpochFeatureBytes, (double)(n_in * row_bytes_));
  p.LogEpochAdd(task->key, kLogEpochMissBytes, (double)(task->num_miss * row_bytes_));
}

// =============================================================================================
// Engine
// =============================================================================================
Engine *Engine::Get() {
  static Engine *e = new Engine();  // leaked on purpose: Python may hold tensors past static destruction
  return e;
}
Engine::Engine() {}
Engine::~Engine() {}

static bool FileExists(const std::string &p) {
  struct stat st;
  return stat(p.c_str(), &st) == 0;
}

void Engine::LoadDataset() {  // engine.cc:73-264
  RunConfig &rc = RunConfig::Get();
  std::string path = rc.dataset_path;
  if (path.empty() || path.back() != '/') path.push_back('/');
  std::unordered_map<std::string, size_t> meta;
  std::ifstream mf(path + "meta.txt");
  FCHECK(mf.good()) << "cannot open " << path << "meta.txt";
  std::string line;
  while (std::getline(mf, line)) {
    std::istringstream iss(line);
    std::vector<std::string> kv{std::istream_iterator<std::string>{iss}, std::istream_iterator<std::string>{}};
    if (kv.size() < 2) break;
    meta[kv[0]] = std::stoull(kv[1]);
  }
  for (const char *k : {"NUM_NODE", "NUM_EDGE", "FEAT_DIM", "NUM_CLASS", "NUM_TRAIN_SET", "NUM_TEST_SET", "NUM_VALID_SET"})
    FCHECK(meta.count(k) > 0) << "meta.txt lacks " << k;
  auto ds = std::make_unique<Dataset>();
  ds->num_node = meta["NUM_NODE"];
  ds->num_edge = meta["NUM_EDGE"];
  ds->num_class = meta["NUM_CLASS"];
  ds->feat_dim = meta["FEAT_DIM"];
  ds->indptr = Tensor::FromMmap(path + "indptr.bin", kI32, {ds->num_node + 1}, "dataset.indptr");
  ds->indices = Tensor::FromMmap(path + "indices.bin", kI32, {ds->num_edge}, "dataset.indices");
  if (FileExists(path + "feat.bin") && rc.option_empty_feat == 0) {
    ds->feat = Tensor::FromMmap(path + "feat.bin", kF32, {ds->num_node, ds->feat_dim}, "dataset.feat");
  } else {
    // engine.cc:144-155: no feature file (or SAMGRAPH_EMPTY_FEAT=k): an uninitialised table, 2^k rows
    const size_t rows = rc.option_empty_feat ? ((size_t)1 << rc.option_empty_feat) : ds->num_node;
    // defer the pinned allocation to the process that owns a GPU (no CUDA before fork)
    ds->feat = std::make_shared<Tensor>();
    ds->feat->dtype = kF32;
    ds->feat->shape = {rows, ds->feat_dim};
    ds->feat->ctx = Context(kCPU, 0);
    ds->feat->nbytes = rows * ds->feat_dim * 4;
    ds->feat->name = "dataset.feat(empty)";
  }
  if (FileExists(path + "label.bin")) {
    ds->label = Tensor::FromMmap(path + "label.bin", kI64, {ds->num_node}, "dataset.label");
  } else {
    ds->label = std::make_shared<Tensor>();
    ds->label->dtype = kI64;
    ds->label->shape = {ds->num_node};
    ds->label->nbytes = ds->num_node * 8;
    ds->label->ctx = Context(kCPU, 0);
  }
  ds->train_set = Tensor::FromMmap(path + "train_set.bin", kI32, {meta["NUM_TRAIN_SET"]}, "dataset.train_set");
  ds->test_set = Tensor::FromMmap(path + "test_set.bin", kI32, {meta["NUM_TEST_SET"]}, "dataset.test_set");
  ds->valid_set = Tensor::FromMmap(path + "valid_set.bin", kI32, {meta["NUM_VALID_SET"]}, "dataset.valid_set");
  if (rc.sample_type == kWeightedKHop || rc.sample_type == kWeightedKHopHashDedup) {
    ds->prob_table = Tensor::FromMmap(path + "prob_table.bin", kF32, {ds->num_edge}, "dataset.prob_table");
    ds->alias_table = Tensor::FromMmap(path + "alias_table.bin", kI32, {ds->num_edge}, "dataset.alias_table");
  } else if (rc.sample_type == kWeightedKHopPrefix) {
    ds->prob_prefix_table = Tensor::FromMmap(path + "prob_prefix_table.bin", kF32, {ds->num_edge}, "dataset.prefix");
  }
  if (rc.UseGPUCache()) {  // engine.cc:216-256
    const char *f = nullptr;
    switch (rc.cache_policy) {
      case kCacheByDegree: f = "cache_by_degree.bin"; break;
      case kCacheByHeuristic: f = "cache_by_heuristic.bin"; break;
      case kCacheByDegreeHop: f = "cache_by_degree_hop.bin"; break;
      case kCacheByFakeOptimal: f = "cache_by_fake_optimal.bin"; break;
      case kCacheByRandom: f = "cache_by_random.bin"; break;
      case kCacheByPreSample: break;
      default: FCHECK(false) << "cache policy " << rc.cache_policy << " is outside the hot-path scope";
    }
    // cache_by_degree / cache_by_random / cache_by_heuristic: when the offline tool's file is absent the ranking is computed on the
    // sampler GPU at init (Engine::DoGpuRanking); degree_hop / fake_optimal are then built here with the host
    // restatement of the reference's offline tools (fgnn_dataset_tools.h; no CUDA: this runs before the fork)
    const bool on_gpu_ok = rc.cache_policy == kCacheByDegree || rc.cache_policy == kCacheByRandom ||
                           rc.cache_policy == kCacheByHeuristic;
    const bool on_host_ok = rc.cache_policy == kCacheByDegreeHop || rc.cache_policy == kCacheByFakeOptimal;
    if (f && !FileExists(path + f) && on_host_ok) {
      Timer tb;
      auto store = std::make_shared<std::vector<uint32_t>>(ds->num_node);
      auto rank = Tensor::View(store->data(), kI32, {ds->num_node}, Context(kCPU, 0), store, "dataset.ranking_nodes");
      const auto *indptr = (const uint32_t *)ds->indptr->data, *indices = (const uint32_t *)ds->indices->data;
      const auto *train = (const uint32_t *)ds->train_set->data;
      const size_t n_train = ds->train_set->NumItems();
      int rc_build;
      if (rc.cache_policy == kCacheByDegreeHop) {
        rc_build = fgnn_rt_rank_degree_hop(indptr, indices, ds->num_node, train, n_train, 2, (int)rc.omp_thread_num,
                                           (uint32_t *)rank->data);
      } else {  // the tool's constants: fanout {25, 10}, 48 threads (cache_by_fake_optimal.cc:170, options.cc:28)
        rc_build = fgnn_rt_rank_fake_optimal(indptr, indices, ds->num_node, train, n_train, 25, 10, 48,
                                             (int)rc.omp_thread_num, (uint32_t *)rank->data);
      }
      FCHECK_EQ(rc_build, 0) << "building " << f << " failed";
      FLOG(Info) << f << " absent: ranking built on the host in " << tb.Passed() << " s";
      ds->ranking_nodes = rank;
    } else if (f && (FileExists(path + f) || !on_gpu_ok)) {
      ds->ranking_nodes = Tensor::FromMmap(path + f, kI32, {ds->num_node}, "dataset.ranking_nodes");
    }
  }
  dataset_ = std::move(ds);
}

// Miss rows are read by the GPU straight out of pinned host memory: keep that memory (and this process's host
// threads) on the NUMA node the GPU hangs off, or at 8 GPUs half of the trainers pull their misses across the
// socket interconnect (r1 SCALE: 39 GB/s per GPU at N = 1, 24 GB/s at N = 8).  Best effort, silent when sysfs
// does not say; FGNN_NUMA_BIND=0 switches it off.  Called before the pinned allocations (first touch decides).
static void BindToGpuNumaNode(int dev) {
  if (GetEnv("FGNN_NUMA_BIND") == "0") return;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return; }
  for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
  std::ifstream nf(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
  int node = -1;
  if (!(nf >> node) || node < 0) return;
  std::ifstream cf("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
  std::string list;
  if (!std::getline(cf, list) || list.empty()) return;
  cpu_set_t set;
  CPU_ZERO(&set);
  std::stringstream ss(list);
  std::string tok;
  int ncpu = 0;
  while (std::getline(ss, tok, ',')) {
    int a = 0, b = 0;
    if (sscanf(tok.c_str(), "%d-%d", &a, &b) == 2) { for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); ++ncpu; } }
    else if (sscanf(tok.c_str(), "%d", &a) == 1 && a < CPU_SETSIZE) { CPU_SET(a, &set); ++ncpu; }
  }
  if (ncpu == 0) return;
  if (sched_setaffinity(0, sizeof(set), &set) == 0)
    FLOG(Info) << "cuda:" << dev << " (" << bus << ") is on NUMA node " << node << ": process bound to its " << ncpu << " cpus";
}

// materialise feature / label buffers that have no backing file (needs CUDA -> post-fork only)
static void EnsureHostTables(Dataset *ds) {
  if (!ds->feat->data) {
    auto t = Tensor::Pinned(kF32, ds->feat->shape, "dataset.feat(empty)");
    memset(t->data, 0, t->nbytes);
    ds->feat = t;
  }
  if (!ds->label->data) {
    auto t = Tensor::Pinned(kI64, ds->label->shape, "dataset.label(empty)");
    memset(t->data, 0, t->nbytes);
    ds->label = t;
  }
}

void Engine::CreateSharedState() {  // dist_engine.cc:115-153 + memory_queue.cc:33-41
  RunConfig &rc = RunConfig::Get();
  const size_t L = fanout_.size();
  size_t slot = sizeof(SlotHeader) + 256;
  size_t cur = batch_size_;
  for (int i = (int)L - 1; i >= 0; --i) {
    slot += cur * fanout_[i] * 4 * 3 + 256 * 3;
    cur += cur * fanout_[i];
  }
  slot += (cur + batch_size_) * 4 + 512;
  slot = (slot + 4095) & ~(size_t)4095;
  const uint32_t nslots = (uint32_t)std::max<size_t>(2, std::min<size_t>(rc.max_copying_jobs + 1, 16));
  const size_t hdr = (sizeof(SharedRing) + 4095) & ~(size_t)4095;
  const size_t rank_bytes = (dataset_->num_node * 4 + 4095) & ~(size_t)4095;
  shared_bytes_ = hdr + rank_bytes + slot * nslots;
  shared_base_ = mmap(nullptr, shared_bytes_, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  FCHECK(shared_base_ != MAP_FAILED) << "mmap of the shared queue failed";
  ring_ = new (shared_base_) SharedRing();
  RingInit(&ring_->ctl, nslots, /*process_shared=*/true);
  pthread_barrierattr_t ba;
  pthread_barrierattr_init(&ba);
  pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&ring_->sampler_barrier, &ba, (unsigned)rc.num_sample_worker);
  pthread_barrier_init(&ring_->trainer_barrier, &ba, (unsigned)rc.num_train_worker);
  ring_->slot_bytes = slot;
  ring_->ranking_off = hdr;
  ring_->slots_off = hdr + rank_bytes;
  ring_->presample_done = 0;
  ring_->live_workers = 0;
  ring_->left_workers = 0;
  // SAMGRAPH_NVLINK_QUEUE=0 restores the reference's D2H -> pinned slot -> H2D bounce (task_queue.cc:131-137,241-255)
  ring_->devq_enabled = GetEnv("SAMGRAPH_NVLINK_QUEUE") == "0" ? 0u : 1u;
  ring_->devq_trainers = (uint32_t)std::min<size_t>(rc.num_train_worker, 16);
  if (rc.num_train_worker > 16) ring_->devq_enabled = 0;
  for (int t = 0; t < 16; ++t) ring_->devq_ready[t] = 0;
  for (uint32_t i = 0; i < nslots; ++i) RingInitSlot(&reinterpret_cast<SlotHeader *>(ring_->slot(i))->ready, i);
}

void Engine::Init() {
  if (initialized_) return;
  RunConfig &rc = RunConfig::Get();
  FCHECK(rc.is_configured);
  Timer t_init;
  FCHECK(rc.run_arch != kArch0) << "arch0 (CPU sampling) is not provided: this runtime has no CPU fallback";
  FCHECK(rc.run_arch == kArch1 || rc.run_arch == kArch2 || rc.run_arch == kArch3 || rc.run_arch == kArch5)
      << "arch" << rc.run_arch << " is outside the hot-path scope (supported: arch1, arch2, arch3, arch5)";
  batch_size_ = rc.batch_size;
  fanout_ = rc.fanout;
  num_epoch_ = rc.num_epoch;
  Timer t_load;
  LoadDataset();
  const double load_time = t_load.Passed();
  const size_t n_train = dataset_->train_set->NumItems();
  num_step_ = (n_train + batch_size_ - 1) / batch_size_;
  num_local_step_ = num_step_;
  Profiler::Get().Reset(num_epoch_, num_step_);
  Profiler::Get().LogInit(kLogInitL2LoadDataset, load_time);
  Profiler::Get().LogInit(kLogInitL3LoadDatasetMMap, load_time);
  dist_ = (rc.run_arch == kArch5);
  if (dist_) {
    // arch5: NO CUDA call before fork (dist_engine.cc:611-632); only shared state is created here
    Timer tq;
    CreateSharedState();
    Profiler::Get().LogInit(kLogInitL2DistQueue, tq.Passed());
    Profiler::Get().LogInit(kLogInitL1Common, t_init.Passed());
    return;
  }
  // ---- single process: sampler + extractor in this process (cuda_engine.cc:64-196) ----
  sampler_ctx_ = rc.sampler_ctx;
  trainer_ctx_ = rc.trainer_ctx;
  role_ = kRoleBoth;
  if (trainer_ctx_.device_type == kGPU) BindToGpuNumaNode(trainer_ctx_.device_id);
  EnsureHostTables(dataset_.get());
  Timer t_state;
  sampler_.reset(new Sampler(dataset_.get(), sampler_ctx_, 0, 1, num_epoch_));
  graph_pool_.reset(new TaskPool(rc.max_copying_jobs));
  Profiler::Get().LogInit(kLogInitL2InternalState, t_state.Passed());
  TensorPtr rank_dev;
  if (rc.UseGPUCache() && rc.cache_policy == kCacheByPreSample) {
    Timer tp;
    DoPreSample();
    Profiler::Get().LogInit(kLogInitL2Presample, tp.Passed());
  } else if (rc.UseGPUCache() && !dataset_->ranking_nodes) {
    DoGpuRanking();
  }
  Timer tc;
  const IdType *rank_host = dataset_->ranking_nodes ? (const IdType *)dataset_->ranking_nodes->data : nullptr;
  extractor_.reset(new Extractor(dataset_.get(), trainer_ctx_, rank_host, nullptr, 0, 0, 1, nullptr));
  Profiler::Get().LogInit(kLogInitL2BuildCache, tc.Passed());
  Profiler::Get().LogInit(kLogInitL1Common, t_init.Passed());
  initialized_ = true;
}

// PreSampler::DoPreSample + GetRankNode (cuda/pre_sampler.cc:57-142): presample_epoch epochs of the
// real sampler, hotness counted in HBM, ranking by one descending radix sort.
void Engine::DoPreSample() {
  RunConfig &rc = RunConfig::Get();
  const size_t V = dataset_->num_node;
  Sampler *s = sampler_.get();
  CUDA_CALL(cudaSetDevice(s->device()));
  auto freq = Tensor::Device(kI32, {V}, s->device(), s->stream(), "presc_freq");
  CUDA_CALL(cudaMemsetAsync(freq->data, 0, V * 4, s->stream()));
  Timer ts;
  // a temporary sampler view limited to presample_epoch epochs: reuse Next() by bounding the loop
  const size_t total = (size_t)std::max(1, rc.presample_epoch) * s->NumLocalStep();
  CUDA_CALL(cudaStreamSynchronize(s->stream()));  // freq is zeroed before any slot stream adds to it
  s->SetRngSalt(0x5052455343000000ull);  // "PRESC": not the stream of the training epochs (see SetRngSalt)
  std::deque<TaskPtr> hold;  // a batch's block may only return to the pool once its slot has drained
  for (size_t i = 0; i < total;) {
    while (hold.size() + s->GroupSize() > s->NumSlots()) {
      s->SyncSlot(hold.front());
      hold.pop_front();
    }
    std::vector<TaskPtr> grp;
    s->NextGroup(&grp);
    while (!grp.empty() && i + grp.size() > total) grp.pop_back();  // drawn past the pre-sampling epochs: not sampled
    if (grp.empty()) break;
    s->EnqueueGroup(grp);  // no count read-back wait: PreSC only needs the unique lists on the device
    for (auto &task : grp) {
      s->CountFrequency(task, (uint32_t *)freq->data);
      hold.push_back(task);
    }
    i += grp.size();
  }
  s->SyncAll();
  hold.clear();
  s->SetRngSalt(0);
  Profiler::Get().LogInit(kLogInitL3PresampleSample, ts.Passed());
  Timer tr;
  auto rank = Tensor::Device(kI32, {V}, s->device(), s->stream(), "presc_rank");
  auto ws = Tensor::Device(kU8, {fgnn_k_presc_rank_workspace_bytes(V)}, s->device(), s->stream(), "presc_ws");
  FGNN_CALL(fgnn_k_presc_rank((const uint32_t *)freq->data, V, (uint32_t *)rank->data, ws->data, ws->nbytes,
                              (fgnn_stream_t)s->stream()));
  IdType *dst_host;
  TensorPtr host_rank;
  if (ring_) {
    dst_host = ring_->ranking();  // published to the trainers through shared memory (dist_engine.cc:119-123)
  } else {
    host_rank = Tensor::Pinned(kI32, {V}, "ranking_nodes");
    dst_host = (IdType *)host_rank->data;
  }
  CUDA_CALL(cudaMemcpyAsync(dst_host, rank->data, V * 4, cudaMemcpyDeviceToHost, s->stream()));
  CUDA_CALL(cudaStreamSynchronize(s->stream()));
  if (host_rank) dataset_->ranking_nodes = host_rank;
  Profiler::Get().LogInit(kLogInitL3PresampleSort, tr.Passed());
  // pre_sampler.cc:101-103: rewind the shuffler and drop the step/epoch logs written while pre-sampling
  s->ResetShuffler();
  Profiler::Get().Reset(num_epoch_, num_step_);
}

// cache_by_degree.cc:36-58 / cache_by_random.cc:36-48 / cache_by_heuristic.cc:28-91 on the sampler GPU, for datasets that ship without the
// offline tool's ranking file: out-degree histogram of the CSR + the PreSC {key,id} descending sort, or a seeded
// permutation.  Published exactly like the PreSC ranking.
void Engine::DoGpuRanking() {
  RunConfig &rc = RunConfig::Get();
  const size_t V = dataset_->num_node, E = dataset_->num_edge;
  Sampler *s = sampler_.get();
  CUDA_CALL(cudaSetDevice(s->device()));
  Timer tr;
  auto rank = Tensor::Device(kI32, {V}, s->device(), s->stream(), "policy_rank");
  if (rc.cache_policy == kCacheByDegree) {
    auto deg = Tensor::Device(kI32, {V}, s->device(), s->stream(), "out_degree");
    auto ws = Tensor::Device(kU8, {fgnn_k_presc_rank_workspace_bytes(V)}, s->device(), s->stream(), "rank_ws");
    FGNN_CALL(fgnn_k_rank_by_degree(s->d_indices(), E, V, (uint32_t *)deg->data, (uint32_t *)rank->data, ws->data,
                                    ws->nbytes, (fgnn_stream_t)s->stream()));
    CUDA_CALL(cudaStreamSynchronize(s->stream()));
  } else if (rc.cache_policy == kCacheByHeuristic) {  // cache_by_heuristic.cc:28-91
    const size_t T = dataset_->train_set->NumItems();
    auto train = Tensor::Device(kI32, {T}, s->device(), s->stream(), "train_set");
    auto total = Tensor::Device(kI64, {1}, s->device(), s->stream(), "num_neighbours");
    CUDA_CALL(cudaMemcpyAsync(train->data, dataset_->train_set->data, T * 4, cudaMemcpyDefault, s->stream()));
    FGNN_CALL(fgnn_k_row_len_sum(s->d_indptr(), (const uint32_t *)train->data, T, (unsigned long long *)total->data,
                                 (fgnn_stream_t)s->stream()));
    unsigned long long n_nbr = 0;
    CUDA_CALL(cudaMemcpyAsync(&n_nbr, total->data, 8, cudaMemcpyDeviceToHost, s->stream()));
    CUDA_CALL(cudaStreamSynchronize(s->stream()));
    auto ws = Tensor::Device(kU8, {fgnn_k_rank_heuristic_workspace_bytes(V, T, (size_t)n_nbr)}, s->device(),
                             s->stream(), "rank_ws");
    FGNN_CALL(fgnn_k_rank_by_heuristic(s->d_indptr(), s->d_indices(), V, E, (const uint32_t *)train->data, T,
                                       (size_t)n_nbr, (uint32_t *)rank->data, ws->data, ws->nbytes,
                                       (fgnn_stream_t)s->stream()));
    CUDA_CALL(cudaStreamSynchronize(s->stream()));
  } else {
    FCHECK(rc.cache_policy == kCacheByRandom) << "cache policy " << rc.cache_policy << " needs its ranking file";
    auto ws = Tensor::Device(kU8, {fgnn_k_rank_random_workspace_bytes(V)}, s->device(), s->stream(), "rank_ws");
    FGNN_CALL(fgnn_k_rank_random(V, rc.seed, (uint32_t *)rank->data, ws->data, ws->nbytes,
                                 (fgnn_stream_t)s->stream()));
    CUDA_CALL(cudaStreamSynchronize(s->stream()));
  }
  IdType *dst_host;
  TensorPtr host_rank;
  if (ring_) {
    dst_host = ring_->ranking();
  } else {
    host_rank = Tensor::Pinned(kI32, {V}, "ranking_nodes");
    dst_host = (IdType *)host_rank->data;
  }
  CUDA_CALL(cudaMemcpyAsync(dst_host, rank->data, V * 4, cudaMemcpyDeviceToHost, s->stream()));
  CUDA_CALL(cudaStreamSynchronize(s->stream()));
  if (host_rank) dataset_->ranking_nodes = host_rank;
  Profiler::Get().LogInit(kLogInitL3PresampleSort, tr.Passed());
}

void Engine::SampleInit(int worker_id, Context ctx) {  // dist_engine.cc:231-364
  FCHECK(dist_ && !initialized_) << "sample_init is only valid in arch5, once per process";
  RunConfig &rc = RunConfig::Get();
  Timer t0;
  role_ = kRoleSampler;
  worker_id_ = worker_id;
  sampler_ctx_ = ctx;
  trainer_ctx_ = ctx;
  CUDA_CALL(cudaSetDevice(ctx.device_id));
  Timer tp;
  CUDA_CALL(cudaHostRegister(shared_base_, shared_bytes_, cudaHostRegisterPortable));  // memory_queue.cc:47-49
  Profiler::Get().LogInit(kLogInitL3DistQueuePin, tp.Passed());
  sampler_.reset(new Sampler(dataset_.get(), ctx, worker_id, (int)rc.num_sample_worker, num_epoch_));
  num_local_step_ = sampler_->NumLocalStep();
  ring_->live_workers.fetch_add(1);
  if (rc.UseGPUCache() && rc.cache_policy == kCacheByPreSample) {
    Timer tps;
    if (worker_id == 0) {  // dist_engine.cc:323-337: sampler 0 pre-samples over the WHOLE train set
      std::unique_ptr<Sampler> full(new Sampler(dataset_.get(), ctx, 0, 1, (size_t)std::max(1, rc.presample_epoch)));
      std::swap(full, sampler_);
      DoPreSample();
      std::swap(full, sampler_);
      ring_->presample_done.store(1, std::memory_order_release);
    }
    pthread_barrier_wait(&ring_->sampler_barrier);
    Profiler::Get().LogInit(kLogInitL2Presample, tps.Passed());
  } else if (dataset_->ranking_nodes && worker_id == 0) {
    memcpy(ring_->ranking(), dataset_->ranking_nodes->data, dataset_->num_node * 4);
    ring_->presample_done.store(1, std::memory_order_release);
  } else if (rc.UseGPUCache() && !dataset_->ranking_nodes) {
    if (worker_id == 0) {
      DoGpuRanking();
      ring_->presample_done.store(1, std::memory_order_release);
    }
    pthread_barrier_wait(&ring_->sampler_barrier);
  }
  Profiler::Get().Reset(num_epoch_, num_step_);
  Profiler::Get().LogInit(kLogInitL1Sampler, t0.Passed());
  initialized_ = true;
}

void Engine::TrainInit(int worker_id, Context ctx) {  // dist_engine.cc:366-465
  FCHECK(dist_ && !initialized_) << "train_init is only valid in arch5, once per process";
  RunConfig &rc = RunConfig::Get();
  Timer t0;
  role_ = kRoleTrainer;
  worker_id_ = worker_id;
  trainer_ctx_ = ctx;
  sampler_ctx_ = ctx;
  CUDA_CALL(cudaSetDevice(ctx.device_id));
  CUDA_CALL(cudaHostRegister(shared_base_, shared_bytes_, cudaHostRegisterPortable));
  ring_->live_workers.fetch_add(1);
  BindToGpuNumaNode(ctx.device_id);
  EnsureHostTables(dataset_.get());
  graph_pool_.reset(new TaskPool(rc.max_copying_jobs));
  const int T = (int)rc.num_train_worker;
  const bool partition = rc.partition_cache && T > 1 && rc.UseGPUCache();
  if (rc.UseGPUCache()) {
    // the ranking is published by sampler 0 (PreSC, file or GPU-computed policy): the scripts order this with a
    // barrier (notify_sampler_ready / wait_for_sampler_ready); a driver that forgets it must not build the cache
    // from a half-written ranking
    Timer tw;
    while (ring_->presample_done.load(std::memory_order_acquire) == 0) {
      FCHECK(!stop_) << "shutdown while waiting for the cache ranking";
      if (tw.Passed() > 600) FCHECK(false) << "no sampler published the cache ranking within 600 s";
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  }
  Timer tc;
  extractor_.reset(new Extractor(dataset_.get(), ctx, rc.UseGPUCache() ? ring_->ranking() : nullptr, nullptr, 0,
                                 partition ? worker_id : 0, partition ? T : 1, ring_));
  Profiler::Get().LogInit(kLogInitL2BuildCache, tc.Passed());
  if (ring_->devq_enabled) {
    const uint32_t Tq = ring_->devq_trainers;
    FCHECK_LT((uint32_t)worker_id, Tq);
    const uint32_t mine = ring_->ctl.num_slots > (uint32_t)worker_id ? (ring_->ctl.num_slots - worker_id + Tq - 1) / Tq : 0;
    devq_base_.assign(Tq, nullptr);
    FGNN_CALL(fgnn_k_shard_alloc((void **)&devq_base_[worker_id], std::max<size_t>(mine, 1) * ring_->slot_bytes));
    FGNN_CALL(fgnn_k_ipc_export(devq_base_[worker_id], ring_->devq_handle[worker_id]));
    devq_own_ = worker_id;
    ring_->devq_ready[worker_id].store(1, std::memory_order_release);
  }
  FLOG(Info) << "arch5 queue transport: "
             << (ring_->devq_enabled ? "device ring (payload slots in trainer HBM, peer copies)"
                                     : "host bounce (pinned shared-memory slots)")
             << ", trainer " << worker_id << " on cuda:" << ctx.device_id;
  // steps this trainer consumes: step % T == worker_id (train_graphsage.py:298)
  num_local_step_ = num_step_ / T + ((size_t)worker_id < num_step_ % T ? 1 : 0);
  Profiler::Get().LogInit(kLogInitL1Trainer, t0.Passed());
  initialized_ = true;
}

// Payload address of ring slot `idx` in trainer HBM; maps the owning trainer's part of the device ring on first
// use (CUDA IPC, peer access enabled lazily: copies to/from it travel over NVLink, or stay inside the GPU when
// sampler and trainer share one).
char *Engine::DevSlot(uint64_t idx) {
  const uint32_t Tq = ring_->devq_trainers;
  if (devq_base_.empty()) devq_base_.assign(Tq, nullptr);
  const uint64_t s = idx % ring_->ctl.num_slots;
  const uint32_t t = (uint32_t)(s % Tq);
  if (!devq_base_[t]) {
    while (ring_->devq_ready[t].load(std::memory_order_acquire) == 0) {
      if (stop_) return nullptr;
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    void *p = nullptr;
    FGNN_CALL(fgnn_k_ipc_open(ring_->devq_handle[t], &p));
    devq_base_[t] = (char *)p;
  }
  return devq_base_[t] + (s / Tq) * ring_->slot_bytes;
}

// ---- arch5 transport: Task <-> record in a queue slot (task_queue.cc:154-347).  Header in pinned shared host
// memory; payload either next to it (host bounce, the reference's path) or in the trainers' HBM (device queue). ----
void Engine::SendTask(const TaskPtr &t) {
  Timer ts;
  uint64_t idx = 0;
  auto ready_of = [this](uint64_t i) { return ring_->ready_of(i); };
  if (!RingBeginWrite(&ring_->ctl, ready_of, &stop_, &idx)) return;  // shutting down
  char *slot = ring_->slot(idx);
  SlotHeader *h = reinterpret_cast<SlotHeader *>(slot);
  CUDA_CALL(cudaSetDevice(sampler_->device()));
  if (!send_stream_) CUDA_CALL(cudaStreamCreateWithFlags(&send_stream_, cudaStreamNonBlocking));
  cudaStream_t st = send_stream_;
  if (t->ready) CUDA_CALL(cudaStreamWaitEvent(st, t->ready, 0));
  const size_t L = t->graphs.size();
  FCHECK_LE(L, (size_t)8);
  h->num_layer = (uint32_t)L;
  h->have_data = (L && t->graphs[0].data) ? 1 : 0;
  h->key = t->key;
  h->input_size = t->input_nodes->NumItems();
  h->output_size = t->output_nodes->NumItems();
  const size_t hdr_bytes = (sizeof(SlotHeader) + 255) & ~(size_t)255;
  char *const payload = ring_->devq_enabled ? DevSlot(idx) : slot;
  if (!payload) return;  // shutting down
  char *p = payload + hdr_bytes;
  double sent_bytes = 0;
  auto put = [&](const TensorPtr &x) {
    if (!x) return;
    FCHECK_LE((size_t)(p - payload) + x->nbytes, (size_t)ring_->slot_bytes) << "task exceeds the queue slot";
    CUDA_CALL(cudaMemcpyAsync(p, x->data, x->nbytes, cudaMemcpyDefault, st));  // D2H, or D2D into the trainer's HBM
    p += (x->nbytes + 255) & ~(size_t)255;
    sent_bytes += (double)x->nbytes;
  };
  put(t->input_nodes);
  put(t->output_nodes);
  for (size_t i = 0; i < L; ++i) {
    h->num_src[i] = t->graphs[i].num_src;
    h->num_dst[i] = t->graphs[i].num_dst;
    h->num_edge[i] = t->graphs[i].num_edge;
    put(t->graphs[i].row);
    put(t->graphs[i].col);
    put(t->graphs[i].data);
  }
  // The slot is published as soon as the copies' event has fired.  The sampler waits for it here (send_stream_
  // carries nothing but these copies, the slots' sampling streams keep running): a batch must never stay
  // unpublished behind the last sample_once of an epoch share, the scripts wait on a barrier after it.
  cudaEvent_t ev = XferEvent();
  CUDA_CALL(cudaEventRecord(ev, st));
  pending_sends_.push_back(PendingXfer{idx, ev, t});
  PollSends(true);
  Profiler::Get().LogStep(t->key, kLogL1GraphBytes, sent_bytes);
  Profiler::Get().LogStep(t->key, kLogL1SendTime, ts.Passed());
  Profiler::Get().LogEpochAdd(t->key, kLogEpochSampleSendTime, ts.Passed());
}

cudaEvent_t Engine::XferEvent() {
  if (!xfer_events_.empty()) {
    cudaEvent_t e = xfer_events_.back();
    xfer_events_.pop_back();
    return e;
  }
  cudaEvent_t e;
  CUDA_CALL(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return e;
}

void Engine::PollSends(bool drain) {
  while (!pending_sends_.empty()) {
    PendingXfer &x = pending_sends_.front();
    if (drain) CUDA_CALL(cudaEventSynchronize(x.ev));
    else if (cudaEventQuery(x.ev) != cudaSuccess) { cudaGetLastError(); break; }
    SlotHeader *h = reinterpret_cast<SlotHeader *>(ring_->slot(x.idx));
    RingEndWrite(&ring_->ctl, &h->ready, x.idx);
    xfer_events_.push_back(x.ev);
    pending_sends_.pop_front();
  }
}

void Engine::PollRecvs(bool drain) {
  while (!pending_recvs_.empty()) {
    PendingXfer &x = pending_recvs_.front();
    if (drain) CUDA_CALL(cudaEventSynchronize(x.ev));
    else if (cudaEventQuery(x.ev) != cudaSuccess) { cudaGetLastError(); break; }
    SlotHeader *h = reinterpret_cast<SlotHeader *>(ring_->slot(x.idx));
    RingEndRead(&ring_->ctl, &h->ready, x.idx);
    xfer_events_.push_back(x.ev);
    pending_recvs_.pop_front();
  }
}

TaskPtr Engine::RecvTask(bool block) {
  Timer tr;
  uint64_t idx = 0;
  // Never wait on the ring while a slot of ours is still unreleased: with several trainers the ticket we are about
  // to take can map to that very slot (ticket t and t + num_slots), its sampler would wait for our release and we
  // for its publication (seen as a hang of 2 samplers + 6 trainers on 8 x B200 in round 2).  The copies out of the
  // previous slot were enqueued before the previous batch's gather, so this wait is a few tens of microseconds.
  if (extractor_) {
    CUDA_CALL(cudaSetDevice(extractor_->device()));
    PollRecvs(true);
  }
  auto ready_of = [this](uint64_t i) { return ring_->ready_of(i); };
  if (!RingBeginRead(&ring_->ctl, ready_of, &stop_, block, &idx)) return nullptr;
  char *slot = ring_->slot(idx);
  SlotHeader *h = reinterpret_cast<SlotHeader *>(slot);
  const int dev = extractor_->device();
  cudaStream_t st = extractor_->stream();
  CUDA_CALL(cudaSetDevice(dev));
  auto task = std::make_shared<Task>();
  task->key = h->key;
  const char *payload = ring_->devq_enabled ? DevSlot(idx) : slot;
  if (!payload) return nullptr;
  const char *p = payload + ((sizeof(SlotHeader) + 255) & ~(size_t)255);
  auto get = [&](size_t n, const char *name) {
    auto t = Tensor::Device(kI32, {n}, dev, st, name);
    CUDA_CALL(cudaMemcpyAsync(t->data, p, n * 4, cudaMemcpyDefault, st));  // H2D, or D2D out of the device ring
    p += (n * 4 + 255) & ~(size_t)255;
    return t;
  };
  task->input_nodes = get(h->input_size, "input_nodes");
  task->output_nodes = get(h->output_size, "output_nodes");
  task->graphs.resize(h->num_layer);
  for (uint32_t i = 0; i < h->num_layer; ++i) {
    TrainGraph &g = task->graphs[i];
    g.num_src = h->num_src[i];
    g.num_dst = h->num_dst[i];
    g.num_edge = h->num_edge[i];
    g.row = get(g.num_edge, "train_graph.row");
    g.col = get(g.num_edge, "train_graph.col");
    if (h->have_data) g.data = get(g.num_edge, "train_graph.data");
  }
  // no host wait: the extraction kernels are ordered behind these copies on the same stream; the slot goes back
  // to the samplers (PollRecvs) once the copies' event has fired
  cudaEvent_t ev = XferEvent();
  CUDA_CALL(cudaEventRecord(ev, st));
  pending_recvs_.push_back(PendingXfer{idx, ev, nullptr});
  PollRecvs(false);
  Profiler::Get().LogStep(task->key, kLogL1RecvTime, tr.Passed());
  return task;
}

// ---------------------------------------------------------------------------------------------
// The loops.  The reference runs a sampler thread and a data-copy thread that each block on their own
// stream (cuda_engine.cc:198-226, cuda_loops_arch3.cc:54-172).  Here ONE host thread pumps an event-driven
// pipeline: it keeps every sampling slot and the extraction stream fed and only ever *polls* CUDA events,
// so all host work (≈16 launches per batch) overlaps the GPU work of the batches in flight and no two host
// threads contend for the driver (measured r1_f: two blocking threads + 3 slots ran 0.39-0.55 ms/step, the
// same work pumped by one thread 0.26 ms/step).
//   stage 1  fill the sampler slots           Sampler::Next + Enqueue        (no wait)
//   stage 2  sampled -> extraction stream     Sampler::Finish + Extractor::Enqueue   when the slot's event fired
//   stage 3  extracted -> graph pool          Extractor::Finish + Submit     when the batch's event fired
// Batches leave in the order they were drawn.
// ---------------------------------------------------------------------------------------------
static inline void CpuRelax(int &idle) {
  if (++idle < 2000) std::this_thread::yield();
  else std::this_thread::sleep_for(std::chrono::microseconds(20));
}

void Engine::FillSamplerSlots() {
  while (inflight_.size() + sampler_->GroupSize() <= sampler_->NumSlots()) {  // a whole slot group is free
    Timer te;
    std::vector<TaskPtr> grp;
    sampler_->NextGroup(&grp);
    if (grp.empty()) break;
    sampler_->EnqueueGroup(grp);
    const double each = te.Passed() / (double)grp.size();
    for (auto &next : grp) {
      Profiler::Get().LogStep(next->key, kLogL2ShuffleTime, each);  // host time spent enqueueing
      inflight_.push_back(next);
    }
  }
}

void Engine::LogSampled(const TaskPtr &task, double finish_time) {
  auto &p = Profiler::Get();
  const double enqueue_time = p.GetLogStepValue(task->key, kLogL2ShuffleTime);
  p.LogStep(task->key, kLogL1SampleTime, enqueue_time + finish_time);
  p.LogStep(task->key, kLogL2CoreSampleTime, finish_time);
  p.LogEpochAdd(task->key, kLogEpochSampleTime, enqueue_time + finish_time);
}

void Engine::LogExtracted(const TaskPtr &task, double finish_time) {
  auto &p = Profiler::Get();
  const double graph_copy = p.GetLogStepValue(task->key, kLogL2GraphCopyTime);
  const double recv = dist_ ? p.GetLogStepValue(task->key, kLogL1RecvTime) : 0.0;
  p.LogStep(task->key, kLogL1CopyTime, recv + graph_copy + finish_time);
  p.LogStep(task->key, kLogL2CacheCopyTime, finish_time);
  p.LogEpochAdd(task->key, kLogEpochCopyTime, recv + graph_copy + finish_time);
}

// single-process archs: one non-blocking turn of the pump; *delivered is set when a batch reached the graph pool
bool Engine::PumpBoth(bool *delivered) {
  bool progress = false;
  const size_t before = inflight_.size();
  FillSamplerSlots();
  progress |= inflight_.size() != before;
  while (!inflight_.empty() && (int)x_inflight_.size() < Extractor::kDepth && sampler_->Done(inflight_.front())) {
    TaskPtr task = inflight_.front();
    inflight_.pop_front();
    Timer t1;
    sampler_->Finish(task);  // the event has fired: no wait
    LogSampled(task, t1.Passed());
    Timer t0;
    task = extractor_->MoveToTrainer(task, sampler_->device());
    Profiler::Get().LogStep(task->key, kLogL2GraphCopyTime, t0.Passed());
    task->t_extract = Timer::NowMicro();
    extractor_->Enqueue(task);
    x_inflight_.push_back(task);
    progress = true;
  }
  if (!x_inflight_.empty() && !graph_pool_->Full() && extractor_->Done(x_inflight_.front())) {
    TaskPtr task = x_inflight_.front();
    x_inflight_.pop_front();
    extractor_->Finish(task);
    LogExtracted(task, (double)(Timer::NowMicro() - task->t_extract) * 1e-6);
    graph_pool_->Submit(task);
    if (delivered) *delivered = true;
    progress = true;
  }
  return progress;
}

// arch5 sampler process: complete the oldest batch and send it through the shared queue
bool Engine::PumpSampler() {
  Timer t0;
  FillSamplerSlots();
  if (inflight_.empty()) return false;
  TaskPtr task = inflight_.front();
  inflight_.pop_front();
  Timer t1;
  sampler_->Finish(task);
  LogSampled(task, t1.Passed());
  SendTask(task);
  Profiler::Get().LogEpochAdd(task->key, kLogEpochSampleTotalTime, t0.Passed());
  return true;
}

// arch5 trainer process: `depth` batches may be taken from the shared queue ahead of the one being delivered.
// Trainers share ONE queue and each consumes a fixed number of batches, so a trainer must never hold more
// batches than it will deliver: depth is 1 under sample_once and bounded by the remaining count under
// extract_start.
bool Engine::PumpTrainer(int depth) {
  while ((int)x_inflight_.size() < depth) {
    TaskPtr task = RecvTask(/*block=*/x_inflight_.empty());
    if (!task) break;
    task->t_extract = Timer::NowMicro();
    extractor_->Enqueue(task);
    x_inflight_.push_back(task);
  }
  if (x_inflight_.empty()) return false;
  TaskPtr task = x_inflight_.front();
  x_inflight_.pop_front();
  extractor_->Finish(task);
  PollRecvs(true);  // the batch is extracted, so its copies out of the ring are done: the slots go back
  LogExtracted(task, (double)(Timer::NowMicro() - task->t_extract) * 1e-6);
  graph_pool_->Submit(task);
  return true;
}

void Engine::RunSampleOnce() {  // Engine::RunSampleOnce: RunArch3LoopsOnce / RunArch5LoopsOnce
  FCHECK(initialized_);
  if (role_ == kRoleBoth) {
    // exactly one batch reaches the graph pool per call; the slots stay primed across calls
    bool delivered = false;
    int idle = 0;
    while (!delivered && !stop_) {
      if (PumpBoth(&delivered)) { idle = 0; continue; }
      if (inflight_.empty() && x_inflight_.empty()) break;  // all epochs done
      CpuRelax(idle);
    }
  } else if (role_ == kRoleSampler) {
    PumpSampler();
  } else if (role_ == kRoleTrainer) {
    while (!PumpTrainer(1) && !stop_) {}
  }
}

void Engine::Start() {  // GPUEngine::Start, cuda_engine.cc:198-226
  FCHECK(initialized_ && role_ == kRoleBoth) << "start() is for the single-process archs; arch5 uses extract_start";
  threads_.emplace_back([this] {
    int idle = 0;
    while (!stop_) {
      if (PumpBoth(nullptr)) idle = 0;
      else CpuRelax(idle);
    }
  });
}

void Engine::StartExtract(int count) {  // DistEngine::StartExtract, dist_engine.cc:474-483
  FCHECK(initialized_ && role_ == kRoleTrainer);
  threads_.emplace_back([this, count] {
    int left = count;
    while (left > 0 && !stop_) {
      if (graph_pool_->Full()) { std::this_thread::sleep_for(std::chrono::microseconds(1)); continue; }
      if (PumpTrainer(std::min(left, (int)Extractor::kDepth))) --left;
    }
  });
}

TaskPtr Engine::NextBatch() {  // samgraph_get_next_batch, operation.cc:209-221
  FCHECK(initialized_ && graph_pool_);
  current_.reset();
  current_ = graph_pool_->Get(&stop_);
  FCHECK(current_) << "get_next_batch after shutdown";
  return current_;
}

void Engine::Shutdown() {
  if (stop_.exchange(true)) return;
  for (auto &t : threads_)
    if (t.joinable()) t.join();
  threads_.clear();
  current_.reset();
  if (sampler_) sampler_->SyncAll();
  if (ring_ && role_ == kRoleSampler && send_stream_) { cudaSetDevice(sampler_->device()); PollSends(true); }
  if (ring_ && role_ == kRoleTrainer && extractor_) { cudaSetDevice(extractor_->device()); PollRecvs(true); }
  inflight_.clear();
  x_inflight_.clear();
  if (extractor_) { cudaSetDevice(extractor_->device()); cudaStreamSynchronize(extractor_->stream()); }
  if (dist_ && ring_ && initialized_ && role_ != kRoleBoth) {
    // Cache shards and device-ring slots are mapped by the other worker processes (CUDA IPC): nobody unmaps or
    // frees until every initialised worker has stopped touching them.  Bounded: a crashed peer must not hang us.
    // Samplers only hold mappings (closing them is always safe) and just report; trainers own memory and wait.
    ring_->left_workers.fetch_add(1);
    Timer tw;
    while (role_ == kRoleTrainer && ring_->left_workers.load() < ring_->live_workers.load() && tw.Passed() < 10.0)
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
  }
  for (int t = 0; t < (int)devq_base_.size(); ++t)
    if (devq_base_[t] && t != devq_own_) { fgnn_k_ipc_close(devq_base_[t]); devq_base_[t] = nullptr; }
  // the own part of the device ring is left to process exit: a sampler may still be copying into it
  cudaGetLastError();
}

}  // namespace rt
}  // namespace fgnn
