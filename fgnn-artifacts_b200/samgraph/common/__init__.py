"""ctypes front end of the samgraph_* C-ABI (include/samgraph_operation.h).

Mirrors the public surface of the reference's samgraph/common/__init__.py: the enum constants the
training scripts read (kKHop2, kArch5, kLogEpochSampleTime, ...), `sample_types`, `builtin_archs`,
`cache_policies`, `cpu()/gpu()` and the `SamGraphBasics` wrapper.  Enumerator order is defined by
the C++ side (csrc/runtime/rt_common.h, rt_profiler.h) and by the reference's enums
(common.h:38-92, profiler.h:30-160)."""
import ctypes
import os
import sysconfig

# --- enums -------------------------------------------------------------------------------------
def _enum(names, ns=globals()):
    for value, name in enumerate(names):
        ns[name] = value
    return len(names)


_enum(["kCPU", "kMMAP", "kGPU"])
_enum(["kKHop0", "kKHop1", "kWeightedKHop", "kRandomWalk", "kWeightedKHopPrefix", "kKHop2",
       "kWeightedKHopHashDedup"])
_enum(["kArch%d" % i for i in range(8)])
_enum(["kCacheByDegree", "kCacheByHeuristic", "kCacheByPreSample", "kCacheByDegreeHop",
       "kCacheByPreSampleStatic", "kCacheByFakeOptimal", "kDynamicCache", "kCacheByRandom"])

kNumLogInitItems = _enum("""kLogInitL1Common kLogInitL1Sampler kLogInitL1Trainer kLogInitL2LoadDataset
    kLogInitL2DistQueue kLogInitL2Presample kLogInitL2InternalState kLogInitL2BuildCache
    kLogInitL3LoadDatasetMMap kLogInitL3LoadDatasetCopy kLogInitL3DistQueueAlloc kLogInitL3DistQueuePin
    kLogInitL3DistQueuePush kLogInitL3PresampleInit kLogInitL3PresampleSample kLogInitL3PresampleCopy
    kLogInitL3PresampleCount kLogInitL3PresampleSort kLogInitL3PresampleReset kLogInitL3PresampleGetRank
    kLogInitL3InternalStateCreateCtx kLogInitL3InternalStateCreateStream""".split())

kNumLogStepItems = _enum("""kLogL1NumSample kLogL1NumNode kLogL1SampleTime kLogL1SendTime kLogL1RecvTime
    kLogL1CopyTime kLogL1ConvertTime kLogL1TrainTime kLogL1FeatureBytes kLogL1LabelBytes kLogL1IdBytes
    kLogL1GraphBytes kLogL1MissBytes kLogL1PrefetchAdvanced kLogL1GetNeighbourTime
    kLogL2ShuffleTime kLogL2LastLayerTime kLogL2LastLayerSize kLogL2CoreSampleTime kLogL2IdRemapTime
    kLogL2GraphCopyTime kLogL2IdCopyTime kLogL2ExtractTime kLogL2FeatCopyTime kLogL2CacheCopyTime
    kLogL3KHopSampleCooTime kLogL3KHopSampleSortCooTime kLogL3KHopSampleCountEdgeTime
    kLogL3KHopSampleCompactEdgesTime kLogL3RandomWalkSampleCooTime kLogL3RandomWalkTopKTime
    kLogL3RandomWalkTopKStep1Time kLogL3RandomWalkTopKStep2Time kLogL3RandomWalkTopKStep3Time
    kLogL3RandomWalkTopKStep4Time kLogL3RandomWalkTopKStep5Time kLogL3RandomWalkTopKStep6Time
    kLogL3RandomWalkTopKStep7Time kLogL3RandomWalkTopKStep8Time kLogL3RandomWalkTopKStep9Time
    kLogL3RandomWalkTopKStep10Time kLogL3RandomWalkTopKStep11Time kLogL3RemapFillUniqueTime
    kLogL3RemapPopulateTime kLogL3RemapMapNodeTime kLogL3RemapMapEdgeTime kLogL3CacheGetIndexTime
    KLogL3CacheCopyIndexTime kLogL3CacheExtractMissTime kLogL3CacheCopyMissTime
    kLogL3CacheCombineMissTime kLogL3CacheCombineCacheTime""".split())

kNumLogEpochItems = _enum("""kLogEpochSampleTime KLogEpochSampleGetCacheMissIndexTime kLogEpochSampleSendTime
    kLogEpochSampleTotalTime kLogEpochCopyTime kLogEpochConvertTime kLogEpochTrainTime kLogEpochTotalTime
    kLogEpochFeatureBytes kLogEpochMissBytes""".split())

_enum("""kL0Event_Train_Step kL1Event_Sample kL2Event_Sample_Shuffle kL2Event_Sample_Core
    kL2Event_Sample_IdRemap kL1Event_Copy kL2Event_Copy_Id kL2Event_Copy_Graph kL2Event_Copy_Extract
    kL2Event_Copy_FeatCopy kL2Event_Copy_CacheCopy kL3Event_Copy_CacheCopy_GetIndex
    kL3Event_Copy_CacheCopy_CopyIndex kL3Event_Copy_CacheCopy_ExtractMiss kL3Event_Copy_CacheCopy_CopyMiss
    kL3Event_Copy_CacheCopy_CombineMiss kL3Event_Copy_CacheCopy_CombineCache kL1Event_Convert
    kL1Event_Train""".split())


def cpu(device_id=0):
    return "cpu:%d" % device_id


def gpu(device_id=0):
    return "cuda:%d" % device_id


sample_types = dict(khop0=kKHop0, khop1=kKHop1, khop2=kKHop2, random_walk=kRandomWalk,
                    weighted_khop=kWeightedKHop, weighted_khop_prefix=kWeightedKHopPrefix,
                    weighted_khop_hash_dedup=kWeightedKHopHashDedup)

cache_policies = dict(degree=kCacheByDegree, heuristic=kCacheByHeuristic, pre_sample=kCacheByPreSample,
                      degree_hop=kCacheByDegreeHop, presample_static=kCacheByPreSampleStatic,
                      fake_optimal=kCacheByFakeOptimal, dynamic_cache=kDynamicCache, random=kCacheByRandom)

# sampler / trainer placement of the single-process archs (common.h:60-68)
builtin_archs = {"arch%d" % i: {"arch": i} for i in range(8)}
builtin_archs["arch0"].update(sampler_ctx=cpu(), trainer_ctx=gpu(0))
builtin_archs["arch1"].update(sampler_ctx=gpu(0), trainer_ctx=gpu(0))
builtin_archs["arch2"].update(sampler_ctx=gpu(0), trainer_ctx=gpu(0))
builtin_archs["arch3"].update(sampler_ctx=gpu(0), trainer_ctx=gpu(1))
builtin_archs["arch4"].update(sampler_ctx=gpu(1), trainer_ctx=gpu(0))

# --- C-ABI table: name -> (argtypes, restype) -----------------------------------------------------
_u64, _int, _dbl, _sz, _str = ctypes.c_uint64, ctypes.c_int, ctypes.c_double, ctypes.c_size_t, ctypes.c_char_p
_ABI = {
    "samgraph_init": ((), None), "samgraph_start": ((), None), "samgraph_shutdown": ((), None),
    "samgraph_data_init": ((), None), "samgraph_sample_once": ((), None),
    "samgraph_sample_init": ((_int, _str), None), "samgraph_train_init": ((_int, _str), None),
    "samgraph_switch_init": ((_int, _str, _dbl), None), "samgraph_extract_start": ((_int,), None),
    "samgraph_num_epoch": ((), _sz), "samgraph_steps_per_epoch": ((), _sz), "samgraph_num_class": ((), _sz),
    "samgraph_feat_dim": ((), _sz), "samgraph_num_local_step": ((), _sz),
    "samgraph_get_next_batch": ((), _u64),
    "samgraph_get_graph_num_src": ((_u64, _int), _sz), "samgraph_get_graph_num_dst": ((_u64, _int), _sz),
    "samgraph_get_graph_num_edge": ((_u64, _int), _sz),
    "samgraph_log_step": ((_u64, _u64, _int, _dbl), None), "samgraph_log_step_add": ((_u64, _u64, _int, _dbl), None),
    "samgraph_log_epoch_add": ((_u64, _int, _dbl), None),
    "samgraph_get_log_init_value": ((_int,), _dbl), "samgraph_get_log_step_value": ((_u64, _u64, _int), _dbl),
    "samgraph_get_log_epoch_value": ((_u64, _int), _dbl),
    "samgraph_report_init": ((), None), "samgraph_report_step": ((_u64, _u64), None),
    "samgraph_report_step_average": ((_u64, _u64), None), "samgraph_report_epoch": ((_u64,), None),
    "samgraph_report_epoch_average": ((_u64,), None), "samgraph_report_node_access": ((), None),
    "samgraph_trace_step_begin": ((_u64, _int, _u64), None), "samgraph_trace_step_end": ((_u64, _int, _u64), None),
    "samgraph_trace_step_begin_now": ((_u64, _int), None), "samgraph_trace_step_end_now": ((_u64, _int), None),
    "samgraph_dump_trace": ((), None), "samgraph_forward_barrier": ((), None),
    "samgraph_wait_one_child": ((), _int),
}


def _extension_path(pkg_file, name):
    here = os.path.dirname(os.path.abspath(pkg_file))
    for suffix in (sysconfig.get_config_var("EXT_SUFFIX"), ".so"):
        if suffix and os.path.exists(os.path.join(here, name + suffix)):
            return os.path.join(here, name + suffix)
    raise ImportError("%s/%s.so is missing: build it with `python fgnn-artifacts_b200/build.py` "
                      "(there is no CPU fallback)" % (here, name))


class SamGraphBasics(object):
    """Thin method-per-entry-point wrapper; ctypes releases the GIL around every call."""

    def __init__(self, pkg_path, *args):
        self.C_LIB_CTYPES = ctypes.CDLL(_extension_path(pkg_path, args[-1]), mode=ctypes.RTLD_GLOBAL)
        for name, (argtypes, restype) in _ABI.items():
            fn = getattr(self.C_LIB_CTYPES, name)
            fn.argtypes, fn.restype = list(argtypes), restype
            short = name[len("samgraph_"):]
            if not hasattr(type(self), short):
                setattr(self, short, fn)

    def config(self, run_config):
        """dict -> two char*[] arrays; list values are space-joined (operation.cc:123-134)."""
        keys = [str(k).encode() for k in run_config]
        vals = [(" ".join(str(x) for x in v) if isinstance(v, (list, tuple)) else str(v)).encode()
                for v in run_config.values()]
        n = len(keys)
        self.C_LIB_CTYPES.samgraph_config.restype = None
        return self.C_LIB_CTYPES.samgraph_config((ctypes.c_char_p * n)(*keys), (ctypes.c_char_p * n)(*vals),
                                                 ctypes.c_size_t(n))

    def sample_init(self, worker_id, ctx):
        return self.C_LIB_CTYPES.samgraph_sample_init(worker_id, str(ctx).encode())

    def train_init(self, worker_id, ctx):
        return self.C_LIB_CTYPES.samgraph_train_init(worker_id, str(ctx).encode())

    def switch_init(self, worker_id, ctx, cache_percentage):
        return self.C_LIB_CTYPES.samgraph_switch_init(worker_id, str(ctx).encode(), cache_percentage)

    def get_graph_num_src(self, key, graph_id):
        return self.C_LIB_CTYPES.samgraph_get_graph_num_src(key, graph_id)

    def get_graph_num_dst(self, key, graph_id):
        return self.C_LIB_CTYPES.samgraph_get_graph_num_dst(key, graph_id)
