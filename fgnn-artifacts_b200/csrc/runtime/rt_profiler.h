// Profiler: the log/trace getters the training scripts read (profiler.h:30-233).
// Enumerator order is part of the Python API (samgraph/common/__init__.py:150-262).
#pragma once
#include <mutex>
#include <ostream>
#include <vector>

#include "rt_common.h"

namespace fgnn {
namespace rt {

enum LogInitItem {
  kLogInitL1Common = 0, kLogInitL1Sampler, kLogInitL1Trainer,
  kLogInitL2LoadDataset, kLogInitL2DistQueue, kLogInitL2Presample, kLogInitL2InternalState, kLogInitL2BuildCache,
  kLogInitL3LoadDatasetMMap, kLogInitL3LoadDatasetCopy, kLogInitL3DistQueueAlloc, kLogInitL3DistQueuePin,
  kLogInitL3DistQueuePush, kLogInitL3PresampleInit, kLogInitL3PresampleSample, kLogInitL3PresampleCopy,
  kLogInitL3PresampleCount, kLogInitL3PresampleSort, kLogInitL3PresampleReset, kLogInitL3PresampleGetRank,
  kLogInitL3InternalStateCreateCtx, kLogInitL3InternalStateCreateStream, kNumLogInitItems
};

enum LogStepItem {
  kLogL1NumSample = 0, kLogL1NumNode, kLogL1SampleTime, kLogL1SendTime, kLogL1RecvTime, kLogL1CopyTime,
  kLogL1ConvertTime, kLogL1TrainTime, kLogL1FeatureBytes, kLogL1LabelBytes, kLogL1IdBytes, kLogL1GraphBytes,
  kLogL1MissBytes, kLogL1PrefetchAdvanced, kLogL1GetNeighbourTime,
  kLogL2ShuffleTime, kLogL2LastLayerTime, kLogL2LastLayerSize, kLogL2CoreSampleTime, kLogL2IdRemapTime,
  kLogL2GraphCopyTime, kLogL2IdCopyTime, kLogL2ExtractTime, kLogL2FeatCopyTime, kLogL2CacheCopyTime,
  kLogL3KHopSampleCooTime, kLogL3KHopSampleSortCooTime, kLogL3KHopSampleCountEdgeTime,
  kLogL3KHopSampleCompactEdgesTime, kLogL3RandomWalkSampleCooTime, kLogL3RandomWalkTopKTime,
  kLogL3RandomWalkTopKStep1Time, kLogL3RandomWalkTopKStep2Time, kLogL3RandomWalkTopKStep3Time,
  kLogL3RandomWalkTopKStep4Time, kLogL3RandomWalkTopKStep5Time, kLogL3RandomWalkTopKStep6Time,
  kLogL3RandomWalkTopKStep7Time, kLogL3RandomWalkTopKStep8Time, kLogL3RandomWalkTopKStep9Time,
  kLogL3RandomWalkTopKStep10Time, kLogL3RandomWalkTopKStep11Time, kLogL3RemapFillUniqueTime,
  kLogL3RemapPopulateTime, kLogL3RemapMapNodeTime, kLogL3RemapMapEdgeTime, kLogL3CacheGetIndexTime,
  KLogL3CacheCopyIndexTime, kLogL3CacheExtractMissTime, kLogL3CacheCopyMissTime, kLogL3CacheCombineMissTime,
  kLogL3CacheCombineCacheTime, kNumLogStepItems
};

enum LogEpochItem {
  kLogEpochSampleTime = 0, KLogEpochSampleGetCacheMissIndexTime, kLogEpochSampleSendTime,
  kLogEpochSampleTotalTime, kLogEpochCopyTime, kLogEpochConvertTime, kLogEpochTrainTime, kLogEpochTotalTime,
  kLogEpochFeatureBytes, kLogEpochMissBytes, kNumLogEpochItems
};

constexpr int kNumTraceItems = 19;  // profiler.h:136-160

class Profiler {
 public:
  static Profiler &Get();
  void Reset(size_t num_epoch, size_t num_step);      // ResetStepEpoch, profiler.cc:80-115
  void LogInit(int item, double v);
  void LogInitAdd(int item, double v);
  void LogStep(uint64_t key, int item, double v);
  void LogStepAdd(uint64_t key, int item, double v);
  void LogEpochAdd(uint64_t key, int item, double v);
  double GetLogInitValue(int item);
  double GetLogStepValue(uint64_t key, int item);
  double GetLogEpochValue(uint64_t epoch, int item);
  void ReportInit();
  void ReportStep(uint64_t epoch, uint64_t step);
  void ReportStepAverage(uint64_t epoch, uint64_t step);
  void ReportEpoch(uint64_t epoch);
  void ReportEpochAverage(uint64_t epoch);
  void TraceStepBegin(uint64_t key, int item, uint64_t us);
  void TraceStepEnd(uint64_t key, int item, uint64_t us);
  void DumpTrace(std::ostream &os);

 private:
  struct LogData {
    std::vector<double> vals;
    std::vector<char> seen;
    double sum = 0;
    size_t cnt = 0;
  };
  void Put(LogData &d, uint64_t key, double v, bool add);
  std::mutex mu_;
  size_t num_step_ = 1, num_epoch_ = 1;
  std::vector<LogData> init_, step_, epoch_;
  struct Trace { uint64_t begin = 0, end = 0; };
  std::vector<std::vector<Trace>> trace_;
};

}  // namespace rt
}  // namespace fgnn
