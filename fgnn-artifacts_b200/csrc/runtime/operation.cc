// The samgraph_* C-ABI (include/samgraph_operation.h), entry for entry the surface of the
// reference's samgraph/common/operation.{h,cc}.  Errors follow the reference: a violated
// FCHECK logs and abort()s the process (logging.cc:69-73); there are no error codes.
#include <signal.h>
#include <sys/wait.h>

#include <iostream>

#include "rt_engine.h"
#include "samgraph_operation.h"

using namespace fgnn::rt;

extern "C" {

void samgraph_config(const char **config_keys, const char **config_values, const size_t num_config_items) {
  RunConfig &RC = RunConfig::Get();  // operation.cc:45-169
  FCHECK(!RC.is_configured);
  std::unordered_map<std::string, std::string> c;
  for (size_t i = 0; i < num_config_items; ++i) c[config_keys[i]] = config_values[i];
  for (const char *k : {"dataset_path", "_arch", "_sample_type", "batch_size", "num_epoch", "_cache_policy",
                        "cache_percentage", "max_sampling_jobs", "max_copying_jobs", "omp_thread_num"})
    FCHECK(c.count(k)) << "missing config key " << k;
  RC.raw = c;
  RC.dataset_path = c["dataset_path"];
  RC.run_arch = (RunArch)std::stoi(c["_arch"]);
  RC.sample_type = (SampleType)std::stoi(c["_sample_type"]);
  RC.batch_size = std::stoull(c["batch_size"]);
  RC.num_epoch = std::stoull(c["num_epoch"]);
  RC.cache_policy = (CachePolicy)std::stoi(c["_cache_policy"]);
  RC.cache_percentage = std::stod(c["cache_percentage"]);
  RC.max_sampling_jobs = std::stoull(c["max_sampling_jobs"]);
  RC.max_copying_jobs = std::stoull(c["max_copying_jobs"]);
  RC.omp_thread_num = std::stoi(c["omp_thread_num"]);
  switch (RC.run_arch) {
    case kArch0: case kArch1: case kArch2: case kArch3: case kArch4:
      FCHECK(c.count("sampler_ctx"));
      FCHECK(c.count("trainer_ctx"));
      RC.sampler_ctx = Context(c["sampler_ctx"]);
      RC.trainer_ctx = Context(c["trainer_ctx"]);
      break;
    case kArch5:
      FCHECK(c.count("num_sample_worker"));
      FCHECK(c.count("num_train_worker"));
      RC.num_sample_worker = std::stoull(c["num_sample_worker"]);
      RC.num_train_worker = std::stoull(c["num_train_worker"]);
      break;
    default:
      FCHECK(false) << "arch" << RC.run_arch << " (SGNN baseline modes) is outside the hot-path scope";
  }
  if (RC.sample_type != kRandomWalk) {
    FCHECK(c.count("num_fanout"));
    FCHECK(c.count("fanout"));
    const size_t nf = std::stoull(c["num_fanout"]);
    std::stringstream ss(c["fanout"]);
    RC.fanout.clear();
    for (size_t i = 0; i < nf; ++i) { size_t f; ss >> f; RC.fanout.push_back(f); }
  } else {
    for (const char *k : {"random_walk_length", "random_walk_restart_prob", "num_random_walk", "num_neighbor", "num_layer"})
      FCHECK(c.count(k)) << "missing config key " << k;
    RC.random_walk_length = std::stoull(c["random_walk_length"]);
    RC.random_walk_restart_prob = std::stod(c["random_walk_restart_prob"]);
    RC.num_random_walk = std::stoull(c["num_random_walk"]);
    RC.num_neighbor = std::stoull(c["num_neighbor"]);
    RC.num_layer = std::stoull(c["num_layer"]);
    RC.fanout = std::vector<size_t>(RC.num_layer, RC.num_neighbor);
  }
  RC.barriered_epoch = c.count("barriered_epoch") ? std::stoi(c["barriered_epoch"]) : 0;
  RC.presample_epoch = c.count("presample_epoch") ? std::stoi(c["presample_epoch"]) : 0;
  if (c.count("seed")) RC.seed = std::stoull(c["seed"]);                    // ours, optional
  if (c.count("partition_cache")) RC.partition_cache = std::stoi(c["partition_cache"]) != 0;
  else if (RC.run_arch == kArch5) RC.partition_cache = true;                // partitioned over trainer GPUs
  if (c.count("replicate_percentage")) RC.replicate_percentage = std::stod(c["replicate_percentage"]);
  else if (!GetEnv("FGNN_REPLICATE_PCT").empty()) RC.replicate_percentage = std::stod(GetEnv("FGNN_REPLICATE_PCT"));
  RC.LoadFromEnv();
  RC.is_configured = true;
}

void samgraph_init() {
  FCHECK(RunConfig::Get().is_configured);
  Engine::Get()->Init();
}
void samgraph_data_init() {
  FCHECK(RunConfig::Get().is_configured);
  Engine::Get()->Init();
}
void samgraph_sample_init(int worker_id, const char *ctx) { Engine::Get()->SampleInit(worker_id, Context(std::string(ctx))); }
void samgraph_train_init(int worker_id, const char *ctx) { Engine::Get()->TrainInit(worker_id, Context(std::string(ctx))); }
void samgraph_switch_init(int, const char *, double) {
  FCHECK(false) << "samgraph_switch_init (dynamic role switching) is outside the hot-path scope";
}
void samgraph_extract_start(int count) { Engine::Get()->StartExtract(count); }
void samgraph_start() {
  FCHECK(Engine::Get()->Initialized() && !Engine::Get()->IsShutdown());
  Engine::Get()->Start();
}
void samgraph_shutdown() { Engine::Get()->Shutdown(); }

size_t samgraph_num_epoch() { return Engine::Get()->NumEpoch(); }
size_t samgraph_steps_per_epoch() { return Engine::Get()->NumStep(); }
size_t samgraph_num_local_step() { return Engine::Get()->NumLocalStep(); }
// test hook (not part of the reference ABI): the cache ranking the loader produced — the policy's file, or the host
// build of degree_hop / fake_optimal when the file is absent; nullptr before data_init or for GPU-built policies
const uint32_t *fgnn_rt_dataset_ranking(size_t *num_nodes) {
  const Dataset *ds = Engine::Get()->GetDataset();
  if (!ds || !ds->ranking_nodes) return nullptr;
  if (num_nodes) *num_nodes = ds->num_node;
  return (const uint32_t *)ds->ranking_nodes->data;
}
size_t samgraph_num_class() { return Engine::Get()->GetDataset()->num_class; }
size_t samgraph_feat_dim() { return Engine::Get()->GetDataset()->feat_dim; }

uint64_t samgraph_get_next_batch() { return Engine::Get()->NextBatch()->key; }
void samgraph_sample_once() { Engine::Get()->RunSampleOnce(); }
// declared by the reference header (operation.h:95-97) but never defined there
void samgraph_sample() { Engine::Get()->RunSampleOnce(); }
void samgraph_extract() { Engine::Get()->RunSampleOnce(); }

static TaskPtr Current(uint64_t key) {
  TaskPtr b = Engine::Get()->CurrentBatch();
  FCHECK(b) << "no current batch: call samgraph_get_next_batch first";
  (void)key;
  return b;
}
size_t samgraph_get_graph_num_src(uint64_t key, int graph_id) { return Current(key)->graphs.at(graph_id).num_src; }
size_t samgraph_get_graph_num_dst(uint64_t key, int graph_id) { return Current(key)->graphs.at(graph_id).num_dst; }
size_t samgraph_get_graph_num_edge(uint64_t key, int graph_id) { return Current(key)->graphs.at(graph_id).num_edge; }

void samgraph_log_step(uint64_t epoch, uint64_t step, int item, double val) {
  FCHECK_LT(item, (int)kNumLogStepItems);
  Profiler::Get().LogStep(Engine::Get()->BatchKey(epoch, step), item, val);
}
void samgraph_log_step_add(uint64_t epoch, uint64_t step, int item, double val) {
  FCHECK_LT(item, (int)kNumLogStepItems);
  Profiler::Get().LogStepAdd(Engine::Get()->BatchKey(epoch, step), item, val);
}
void samgraph_log_epoch_add(uint64_t epoch, int item, double val) {
  FCHECK_LT(item, (int)kNumLogEpochItems);
  Profiler::Get().LogEpochAdd(Engine::Get()->BatchKey(epoch, 0), item, val);
}
double samgraph_get_log_init_value(int item) {
  FCHECK_LT(item, (int)kNumLogInitItems);
  return Profiler::Get().GetLogInitValue(item);
}
double samgraph_get_log_step_value(uint64_t epoch, uint64_t step, int item) {
  FCHECK_LT(item, (int)kNumLogStepItems);
  return Profiler::Get().GetLogStepValue(Engine::Get()->BatchKey(epoch, step), item);
}
double samgraph_get_log_epoch_value(uint64_t epoch, int item) {
  FCHECK_LT(item, (int)kNumLogEpochItems);
  return Profiler::Get().GetLogEpochValue(epoch, item);
}
void samgraph_report_init() { Profiler::Get().ReportInit(); }
void samgraph_report_step(uint64_t epoch, uint64_t step) { Profiler::Get().ReportStep(epoch, step); }
void samgraph_report_step_average(uint64_t epoch, uint64_t step) { Profiler::Get().ReportStepAverage(epoch, step); }
void samgraph_report_epoch(uint64_t epoch) { Profiler::Get().ReportEpoch(epoch); }
void samgraph_report_epoch_average(uint64_t epoch) { Profiler::Get().ReportEpochAverage(epoch); }
void samgraph_report_node_access() {}  // node-access similarity reports: offline paper analysis, out of scope
void samgraph_trace_step_begin(uint64_t key, int item, uint64_t us) { Profiler::Get().TraceStepBegin(key, item, us); }
void samgraph_trace_step_end(uint64_t key, int item, uint64_t us) { Profiler::Get().TraceStepEnd(key, item, us); }
void samgraph_trace_step_begin_now(uint64_t key, int item) { Profiler::Get().TraceStepBegin(key, item, Timer::NowMicro()); }
void samgraph_trace_step_end_now(uint64_t key, int item) { Profiler::Get().TraceStepEnd(key, item, Timer::NowMicro()); }
void samgraph_dump_trace() { Profiler::Get().DumpTrace(std::cerr); }
void samgraph_forward_barrier() { Engine::Get()->ForwardBarrier(); }

int samgraph_wait_one_child() {  // operation.cc:374-385
  int child_stat = 0;
  pid_t pid = waitpid(-1, &child_stat, 0);
  if (WEXITSTATUS(child_stat) != 0) {
    FLOG(Error) << "detect a terminated child " << pid << ", status is " << WEXITSTATUS(child_stat);
    return 1;
  } else if (WIFSIGNALED(child_stat) && (WTERMSIG(child_stat) == SIGABRT)) {
    FLOG(Error) << "detect an aborted child " << pid;
    return 1;
  }
  return 0;
}

}  // extern "C"
