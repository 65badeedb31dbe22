// Engine: the sampler / extractor / trainer-facing runtime behind the samgraph_*
// C-ABI (reference: engine.h, cuda/cuda_engine.*, cuda/cuda_loops*.cc,
// dist/dist_engine.*, dist/dist_loops*.cc, graph_pool.*, task_queue.*,
// memory_queue.*).  One class covers the single-process archs (arch1/2/3: sampler
// and extractor threads in this process) and the factored multi-process arch5
// (sampler processes and trainer processes forked after data_init, connected by a
// pinned shared-memory queue).
#pragma once
#include <pthread.h>
#include <semaphore.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "rt_common.h"
#include "rt_profiler.h"
#include "rt_ring.h"

extern "C" void fgnn_rt_step_split(size_t num_step, size_t num_worker, size_t worker_id, size_t *begin,
                                   size_t *count);

namespace fgnn {
namespace rt {

// Bounded in-process queue (TaskQueue / GraphPool, task_queue.h, graph_pool.cc:31-57).
class TaskPool {
 public:
  explicit TaskPool(size_t max_size) : max_(max_size ? max_size : 1) {}
  bool Full();
  void Submit(TaskPtr t);
  TaskPtr Get(std::atomic<bool> *stop);   // blocks (1 us polls, like graph_pool.cc:31-49)
  TaskPtr TryGet();
 private:
  std::mutex mu_;
  std::deque<TaskPtr> q_;
  size_t max_;
};

// Process-shared state of arch5, placed in MAP_SHARED memory before fork()
// (dist_engine.cc:115-153, memory_queue.h:46-113).
struct SharedRing;

class Sampler;    // sampler-GPU state + DoGPUSample
class Extractor;  // trainer-GPU state + DoCacheFeatureCopy / label extract

class Engine {
 public:
  static Engine *Get();
  ~Engine();

  // ---- lifecycle (operation.cc:171-186, 335-360) ----
  void Init();                                  // samgraph_init / samgraph_data_init
  void SampleInit(int worker_id, Context ctx);  // samgraph_sample_init (arch5)
  void TrainInit(int worker_id, Context ctx);   // samgraph_train_init (arch5)
  void Start();                                 // samgraph_start: background sampler+extractor threads
  void StartExtract(int count);                 // samgraph_extract_start (arch5 trainer)
  void RunSampleOnce();                         // samgraph_sample_once
  void Shutdown();

  // ---- queries ----
  size_t NumEpoch() const { return num_epoch_; }
  size_t NumStep() const { return num_step_; }
  size_t NumLocalStep() const { return num_local_step_; }
  uint64_t BatchKey(uint64_t epoch, uint64_t step) const { return epoch * num_step_ + step; }
  const Dataset *GetDataset() const { return dataset_.get(); }
  Context SamplerCtx() const { return sampler_ctx_; }
  Context TrainerCtx() const { return trainer_ctx_; }
  bool Initialized() const { return initialized_; }
  bool IsShutdown() const { return stop_; }

  TaskPtr NextBatch();                          // samgraph_get_next_batch
  TaskPtr CurrentBatch() { return current_; }
  void ForwardBarrier() { ++outer_counter_; }

 private:
  Engine();
  void LoadDataset();
  // event-driven pump (see rt_engine.cc "The loops")
  void FillSamplerSlots();
  bool PumpBoth(bool *delivered);
  bool PumpSampler();
  bool PumpTrainer(int depth);
  void LogSampled(const TaskPtr &t, double finish_time);
  void LogExtracted(const TaskPtr &t, double finish_time);
  TaskPtr RecvTask(bool block = true);
  void SendTask(const TaskPtr &t);
  // arch5 transport without a host wait per batch: copies are enqueued, the ring slot is published / released
  // when their event has fired (polled by the pump)
  struct PendingXfer { uint64_t idx; cudaEvent_t ev; TaskPtr keep; };
  void PollSends(bool drain);
  void PollRecvs(bool drain);
  cudaEvent_t XferEvent();
  std::deque<PendingXfer> pending_sends_, pending_recvs_;
  std::vector<cudaEvent_t> xfer_events_;
  void DoPreSample();
  void DoGpuRanking();
  void CreateSharedState();
  char *DevSlot(uint64_t idx);

  bool initialized_ = false;
  std::atomic<bool> stop_{false};
  bool dist_ = false;  // arch5
  enum Role { kRoleBoth, kRoleSampler, kRoleTrainer, kRoleNone } role_ = kRoleNone;
  int worker_id_ = 0;

  std::unique_ptr<Dataset> dataset_;
  Context sampler_ctx_, trainer_ctx_;
  size_t num_epoch_ = 0, num_step_ = 0, num_local_step_ = 0, batch_size_ = 0;
  std::vector<size_t> fanout_;

  std::unique_ptr<Sampler> sampler_;
  std::unique_ptr<Extractor> extractor_;
  std::unique_ptr<TaskPool> graph_pool_; // extractor -> python
  TaskPtr current_;
  std::deque<TaskPtr> inflight_;  // enqueued on a sampler slot, not yet completed (sampler thread only)
  cudaStream_t send_stream_ = nullptr;
  std::deque<TaskPtr> x_inflight_;  // enqueued on the extraction stream, not yet completed (extract thread only)
  std::vector<std::thread> threads_;
  std::atomic<uint64_t> outer_counter_{0};

  SharedRing *ring_ = nullptr;  // arch5
  std::vector<char *> devq_base_;  // device queue: base of every trainer's part of the ring (own or IPC-mapped)
  int devq_own_ = -1;
  void *shared_base_ = nullptr;
  size_t shared_bytes_ = 0;
};

}  // namespace rt
}  // namespace fgnn
