"""CPU tests: the C-ABI libraries load without a GPU and export every symbol the headers declare; the
Python mirror of the reference API keeps the reference's constants."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fgnn-artifacts_b200")


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:fgnn_k|samgraph)_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    klib, clib = os.path.join(PKG, "lib", "libfgnn_kernels.so"), os.path.join(PKG, "samgraph", "torch", "c_lib.so")
    if not (os.path.exists(klib) and os.path.exists(clib)):
        g.build()
    return klib, clib


def test_kernel_library_exports_every_declared_symbol(built):
    names = declared("fgnn_kernels.h")
    assert len(names) >= 30
    for path in built:               # the runtime .so embeds the kernel layer too
        lib = ctypes.CDLL(path)
        for n in names:
            assert hasattr(lib, n), "%s does not export %s" % (path, n)
    lib = ctypes.CDLL(built[0])
    lib.fgnn_k_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.fgnn_k_version()
    lib.fgnn_k_ht_capacity.restype = ctypes.c_size_t
    lib.fgnn_k_ht_capacity.argtypes = [ctypes.c_size_t]
    cap = lib.fgnn_k_ht_capacity(2288000)
    assert cap & (cap - 1) == 0 and cap >= 1.5 * 2288000


def test_runtime_library_exports_reference_abi(built):
    names = declared("samgraph_operation.h")
    # every function of the reference's operation.h:29-108 (+ get_log_init_value, operation.cc:267)
    for must in ("samgraph_config samgraph_init samgraph_start samgraph_shutdown samgraph_num_epoch "
                 "samgraph_steps_per_epoch samgraph_num_class samgraph_feat_dim samgraph_get_next_batch "
                 "samgraph_sample_once samgraph_get_graph_num_src samgraph_get_graph_num_dst "
                 "samgraph_get_graph_num_edge samgraph_log_step samgraph_log_step_add samgraph_log_epoch_add "
                 "samgraph_get_log_init_value samgraph_get_log_step_value samgraph_get_log_epoch_value "
                 "samgraph_report_init samgraph_report_step samgraph_report_step_average samgraph_report_epoch "
                 "samgraph_report_epoch_average samgraph_report_node_access samgraph_trace_step_begin "
                 "samgraph_trace_step_end samgraph_trace_step_begin_now samgraph_trace_step_end_now "
                 "samgraph_dump_trace samgraph_forward_barrier samgraph_data_init samgraph_sample_init "
                 "samgraph_train_init samgraph_sample samgraph_extract samgraph_extract_start "
                 "samgraph_switch_init samgraph_num_local_step samgraph_wait_one_child").split():
        assert must in names
    lib = ctypes.CDLL(built[1])
    for n in names:
        assert hasattr(lib, n), "c_lib.so does not export %s" % n
    assert hasattr(lib, "PyInit_c_lib")


def test_runtime_library_exports_the_dataset_tools(built):
    src = open(os.path.join(ROOT, "include", "fgnn_dataset_tools.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(fgnn_rt_[a-z0-9_]+)\s*\(", src)))
    assert names == ["fgnn_rt_rank_degree_hop", "fgnn_rt_rank_fake_optimal"]
    lib = ctypes.CDLL(built[1])
    for n in names:
        assert hasattr(lib, n), "c_lib.so does not export %s" % n


def test_python_binding_table_covers_the_header(built):
    from fgnn_b200 import kernels
    assert sorted(kernels.exported_symbols()) == declared("fgnn_kernels.h")
    kernels.load()


def test_samgraph_python_api_keeps_reference_constants(built):
    import samgraph.common as sc
    # values fixed by the reference enums (common.h:38-92, profiler.h:30-160, __init__.py:47-262)
    assert (sc.kKHop0, sc.kKHop1, sc.kWeightedKHop, sc.kRandomWalk, sc.kWeightedKHopPrefix, sc.kKHop2,
            sc.kWeightedKHopHashDedup) == (0, 1, 2, 3, 4, 5, 6)
    assert sc.kArch5 == 5 and sc.kCacheByPreSample == 2 and sc.kCacheByRandom == 7
    assert sc.kLogInitL1Common == 0 and sc.kNumLogInitItems == 22
    assert sc.kLogL1NumSample == 0 and sc.kLogL1CopyTime == 5 and sc.kLogL1MissBytes == 12
    assert sc.kLogL2ShuffleTime == 15 and sc.kLogL3KHopSampleCooTime == 25 and sc.kNumLogStepItems == 52
    assert sc.kLogEpochSampleTime == 0 and sc.kLogEpochCopyTime == 4 and sc.kLogEpochMissBytes == 9
    assert sc.kL0Event_Train_Step == 0 and sc.kL1Event_Train == 18
    assert sc.sample_types["khop2"] == 5 and sc.cache_policies["pre_sample"] == 2
    assert sc.builtin_archs["arch3"] == {"arch": 3, "sampler_ctx": "cuda:0", "trainer_ctx": "cuda:1"}
    assert sc.gpu(3) == "cuda:3" and sc.cpu() == "cpu:0"
    import samgraph.torch as sam
    for fn in ("config init start shutdown sample_once get_next_batch get_dgl_blocks get_dgl_blocks_with_weights "
               "get_graph_feat get_graph_label get_graph_row get_graph_col get_graph_data num_class feat_dim "
               "num_epoch steps_per_epoch data_init sample_init train_init extract_start num_local_step "
               "wait_one_child get_log_epoch_value log_step report_step_average forward_barrier "
               "get_dataset_feat get_dataset_label get_graph_input_nodes get_graph_output_nodes load_subtensor "
               "notify_sampler_ready wait_for_sampler_ready trace_step_begin_now dump_trace").split():
        assert callable(getattr(sam, fn)), fn


def test_step_split_matches_dist_shuffler(built):
    lib = ctypes.CDLL(built[1])
    f = lib.fgnn_rt_step_split
    f.argtypes = [ctypes.c_size_t] * 3 + [ctypes.POINTER(ctypes.c_size_t)] * 2
    for num_step in (1, 7, 151, 152):
        for S in (1, 2, 3, 8):
            covered = []
            for w in range(S):
                b, c = ctypes.c_size_t(), ctypes.c_size_t()
                f(num_step, S, w, ctypes.byref(b), ctypes.byref(c))
                # dist_shuffler.cc:60-83: floor(N/S) steps each, the last sampler takes the remainder
                assert b.value == (num_step // S) * w
                assert c.value == (num_step // S if w < S - 1 else num_step - (num_step // S) * w)
                covered += list(range(b.value, b.value + c.value))
            assert covered == list(range(num_step))


def test_every_sam_name_the_reference_scripts_use_exists(built):
    """tests/golden/sam_api_used.json lists every `sam.<name>` of example/samgraph/**/*.py (generated by
    tests/golden/make_api_fixture.py).  Every name the reference's own package defines must exist here with the same
    kind (callable vs constant vs dict); the three names its stale train_gat.py uses without the package defining
    them (simple_hashtable, parallel_hashtable, report) are absent on both sides."""
    import json
    used = json.load(open(os.path.join(ROOT, "tests", "golden", "sam_api_used.json")))
    import samgraph.torch as sam
    assert len(used) >= 70
    for name, info in used.items():
        if not info["defined_in_reference"]:
            assert not hasattr(sam, name), name
            continue
        assert hasattr(sam, name), "samgraph.torch lacks %s (used by %s)" % (name, info["scripts"][:2])
        obj = getattr(sam, name)
        if name.startswith(("kLog", "KLog", "kL", "kArch", "kKHop", "kCache", "kWeighted", "kRandom", "kDynamic")):
            assert isinstance(obj, int), name
        elif name in ("sample_types", "builtin_archs", "cache_policies"):
            assert isinstance(obj, dict) and obj, name
        else:
            assert callable(obj), name
