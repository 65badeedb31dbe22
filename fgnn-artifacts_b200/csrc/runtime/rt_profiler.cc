#include "rt_profiler.h"

#include <cstdio>

namespace fgnn {
namespace rt {

Profiler &Profiler::Get() {
  static Profiler p;
  return p;
}

void Profiler::Reset(size_t num_epoch, size_t num_step) {
  std::lock_guard<std::mutex> lk(mu_);
  num_epoch_ = num_epoch ? num_epoch : 1;
  num_step_ = num_step ? num_step : 1;
  const size_t nkeys = num_epoch_ * num_step_;
  auto mk = [](size_t n, size_t keys) {
    std::vector<LogData> v(n);
    for (auto &d : v) { d.vals.assign(keys, 0.0); d.seen.assign(keys, 0); }
    return v;
  };
  if (init_.empty()) init_ = mk(kNumLogInitItems, 1);
  step_ = mk(kNumLogStepItems, nkeys);
  epoch_ = mk(kNumLogEpochItems, num_epoch_);
  trace_.assign(kNumTraceItems, std::vector<Trace>(nkeys));
}

void Profiler::Put(LogData &d, uint64_t key, double v, bool add) {  // profiler.cc:117-165
  if (key >= d.vals.size()) return;
  if (add) d.vals[key] += v; else d.vals[key] = v;
  d.sum += v;
  if (!d.seen[key]) { d.seen[key] = 1; ++d.cnt; }
}

void Profiler::LogInit(int item, double v) {
  std::lock_guard<std::mutex> lk(mu_);
  if (init_.empty()) { init_.resize(kNumLogInitItems); for (auto &d : init_) { d.vals.assign(1, 0.0); d.seen.assign(1, 0); } }
  Put(init_[item], 0, v, false);
}
void Profiler::LogInitAdd(int item, double v) {
  std::lock_guard<std::mutex> lk(mu_);
  if (init_.empty()) { init_.resize(kNumLogInitItems); for (auto &d : init_) { d.vals.assign(1, 0.0); d.seen.assign(1, 0); } }
  Put(init_[item], 0, v, true);
}
void Profiler::LogStep(uint64_t key, int item, double v) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!step_.empty()) Put(step_[item], key, v, false);
}
void Profiler::LogStepAdd(uint64_t key, int item, double v) {
  std::lock_guard<std::mutex> lk(mu_);
  if (!step_.empty()) Put(step_[item], key, v, true);
}
void Profiler::LogEpochAdd(uint64_t key, int item, double v) {  // key -> epoch (profiler.cc:156-165)
  std::lock_guard<std::mutex> lk(mu_);
  if (!epoch_.empty()) Put(epoch_[item], key / num_step_, v, true);
}
double Profiler::GetLogInitValue(int item) {
  std::lock_guard<std::mutex> lk(mu_);
  return init_.empty() ? 0.0 : init_[item].vals[0];
}
double Profiler::GetLogStepValue(uint64_t key, int item) {
  std::lock_guard<std::mutex> lk(mu_);
  return (step_.empty() || key >= step_[item].vals.size()) ? 0.0 : step_[item].vals[key];
}
double Profiler::GetLogEpochValue(uint64_t epoch, int item) {
  std::lock_guard<std::mutex> lk(mu_);
  return (epoch_.empty() || epoch >= epoch_[item].vals.size()) ? 0.0 : epoch_[item].vals[epoch];
}

static int ProfileLevel() {  // profiler.cc:182-195 SAMGRAPH_PROFILE_LEVEL
  const std::string v = GetEnv("SAMGRAPH_PROFILE_LEVEL");
  return v == "1" ? 1 : v == "2" ? 2 : v == "3" ? 3 : 0;
}

void Profiler::ReportInit() {
  std::lock_guard<std::mutex> lk(mu_);
  if (init_.empty() || ProfileLevel() < 1) return;
  printf("    [Init Profiler Level 1]\n        L1  init %10.4lf | sampler init %10.4lf | trainer init %.4lf\n",
         init_[kLogInitL1Common].vals[0], init_[kLogInitL1Sampler].vals[0], init_[kLogInitL1Trainer].vals[0]);
  if (ProfileLevel() >= 2)
    printf("    [Init Profiler Level 2]\n        L2  load ds %10.4lf | presample %10.4lf | internal %10.4lf | "
           "build cache %.4lf\n",
           init_[kLogInitL2LoadDataset].vals[0], init_[kLogInitL2Presample].vals[0],
           init_[kLogInitL2InternalState].vals[0], init_[kLogInitL2BuildCache].vals[0]);
  fflush(stdout);
}

void Profiler::ReportStep(uint64_t epoch, uint64_t step) {
  std::lock_guard<std::mutex> lk(mu_);
  if (step_.empty() || ProfileLevel() < 1) return;
  const uint64_t key = epoch * num_step_ + step;
  if (key >= step_[0].vals.size()) return;
  printf("    [Step(%lu:%lu) Profiler Level 1]\n        L1  sample %10.4lf | copy %10.4lf | convert %.4lf | train %.4lf\n"
         "        L1  feature nbytes %.0lf | label nbytes %.0lf | id nbytes %.0lf | graph nbytes %.0lf | miss nbytes %.0lf\n",
         (unsigned long)epoch, (unsigned long)step, step_[kLogL1SampleTime].vals[key], step_[kLogL1CopyTime].vals[key],
         step_[kLogL1ConvertTime].vals[key], step_[kLogL1TrainTime].vals[key], step_[kLogL1FeatureBytes].vals[key],
         step_[kLogL1LabelBytes].vals[key], step_[kLogL1IdBytes].vals[key], step_[kLogL1GraphBytes].vals[key],
         step_[kLogL1MissBytes].vals[key]);
  fflush(stdout);
}

void Profiler::ReportStepAverage(uint64_t epoch, uint64_t step) {
  std::lock_guard<std::mutex> lk(mu_);
  if (step_.empty() || ProfileLevel() < 1) return;
  auto avg = [&](int item) { return step_[item].cnt ? step_[item].sum / step_[item].cnt : 0.0; };
  printf("    [Step Average Profiler Level 1 E%lu S%lu]\n        L1  sample %10.4lf | send %10.4lf | recv %10.4lf | copy "
         "%10.4lf | convert %.4lf | train %.4lf\n        L1  num nodes %.0lf | num samples %.0lf | feature nbytes %.0lf | "
         "miss nbytes %.0lf\n",
         (unsigned long)epoch, (unsigned long)step, avg(kLogL1SampleTime), avg(kLogL1SendTime), avg(kLogL1RecvTime),
         avg(kLogL1CopyTime), avg(kLogL1ConvertTime), avg(kLogL1TrainTime), avg(kLogL1NumNode), avg(kLogL1NumSample),
         avg(kLogL1FeatureBytes), avg(kLogL1MissBytes));
  if (ProfileLevel() >= 2)
    printf("    [Step Average Profiler Level 2]\n        L2  shuffle %.4lf | core sample %.4lf | id remap %.4lf | graph copy "
           "%.4lf | cache feat copy %.4lf\n",
           avg(kLogL2ShuffleTime), avg(kLogL2CoreSampleTime), avg(kLogL2IdRemapTime), avg(kLogL2GraphCopyTime),
           avg(kLogL2CacheCopyTime));
  fflush(stdout);
}

void Profiler::ReportEpoch(uint64_t epoch) {
  std::lock_guard<std::mutex> lk(mu_);
  if (epoch_.empty() || epoch >= epoch_[0].vals.size() || ProfileLevel() < 1) return;
  printf("    [Profile of Epoch %lu]\n        L1  sample %.4lf | copy %.4lf | convert %.4lf | train %.4lf | total %.4lf\n",
         (unsigned long)epoch, epoch_[kLogEpochSampleTime].vals[epoch], epoch_[kLogEpochCopyTime].vals[epoch],
         epoch_[kLogEpochConvertTime].vals[epoch], epoch_[kLogEpochTrainTime].vals[epoch],
         epoch_[kLogEpochTotalTime].vals[epoch]);
  fflush(stdout);
}

void Profiler::ReportEpochAverage(uint64_t epoch) {
  std::lock_guard<std::mutex> lk(mu_);
  if (epoch_.empty() || ProfileLevel() < 1) return;
  auto avg = [&](int item) { return epoch_[item].cnt ? epoch_[item].sum / epoch_[item].cnt : 0.0; };
  printf("    [Profile of Epoch Average %lu]\n        L1  sample %.4lf | copy %.4lf | convert %.4lf | train %.4lf | total %.4lf\n",
         (unsigned long)epoch, avg(kLogEpochSampleTime), avg(kLogEpochCopyTime), avg(kLogEpochConvertTime),
         avg(kLogEpochTrainTime), avg(kLogEpochTotalTime));
  fflush(stdout);
}

void Profiler::TraceStepBegin(uint64_t key, int item, uint64_t us) {
  std::lock_guard<std::mutex> lk(mu_);
  if (item < 0 || item >= (int)trace_.size() || key >= trace_[item].size()) return;
  trace_[item][key].begin = us;
}
void Profiler::TraceStepEnd(uint64_t key, int item, uint64_t us) {
  std::lock_guard<std::mutex> lk(mu_);
  if (item < 0 || item >= (int)trace_.size() || key >= trace_[item].size()) return;
  trace_[item][key].end = us;
}

void Profiler::DumpTrace(std::ostream &os) {  // chrome trace JSON, profiler.cc:295-366
  if (!RunConfig::Get().option_dump_trace) return;
  std::lock_guard<std::mutex> lk(mu_);
  os << "[\n";
  bool first = true;
  for (size_t item = 0; item < trace_.size(); ++item)
    for (size_t key = 0; key < trace_[item].size(); ++key) {
      const Trace &t = trace_[item][key];
      if (!t.begin || t.end < t.begin) continue;
      if (!first) os << ",\n";
      first = false;
      os << "{\"name\":\"item" << item << "\",\"ph\":\"X\",\"pid\":0,\"tid\":" << item << ",\"ts\":" << t.begin
         << ",\"dur\":" << (t.end - t.begin) << ",\"args\":{\"key\":" << key << "}}";
    }
  os << "\n]\n";
}

}  // namespace rt
}  // namespace fgnn
