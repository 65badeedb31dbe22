"""Host-side runtime logic behind the samgraph_* C-ABI that needs no GPU: config parsing, the dataset loader
(meta.txt + *.bin, engine.cc:73-264), step / epoch arithmetic, the profiler log API and the error behaviour
(CHECK failure = log + abort, logging.cc:69-73).  Each scenario runs in its own interpreter because the engine is a
process-wide singleton.  arch5's data_init() touches no CUDA (the reference forks after it), so it runs here."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = textwrap.dedent("""
    import json, os, sys
    import numpy as np
    sys.path.insert(0, os.path.join(%r, "fgnn-artifacts_b200"))
    import samgraph.torch as sam
    import samgraph.common as sc
    cfg = json.load(open(sys.argv[1]))
    out = {}
""" % ROOT)


@pytest.fixture(scope="module")
def dataset(tmp_path_factory, oracle):
    from fgnn_b200.synth import make_dataset_numpy, write_dataset
    ds = make_dataset_numpy((5000, 60000, 16, 5, 700), seed=3)
    path = str(tmp_path_factory.mktemp("ds_host"))
    write_dataset(path, ds, with_weights=True, oracle=oracle)
    ds["path"] = path
    return ds


def config(path, **kw):
    import samgraph.common as sc
    cfg = {"dataset_path": path, "_arch": 5, "arch": "arch5", "_sample_type": sc.sample_types["khop2"],
           "sample_type": "khop2", "batch_size": 128, "num_epoch": 3, "_cache_policy": sc.cache_policies["pre_sample"],
           "cache_policy": "pre_sample", "cache_percentage": 0.2, "max_sampling_jobs": 4, "max_copying_jobs": 2,
           "omp_thread_num": 4, "presample_epoch": 1, "fanout": [5, 10], "num_fanout": 2, "num_layer": 2,
           "num_sample_worker": 2, "num_train_worker": 2}
    cfg.update(kw)
    return cfg


def run(tmp_path, cfg, body):
    cfg_path = os.path.join(str(tmp_path), "cfg.json")
    json.dump(cfg, open(cfg_path, "w"))
    script = os.path.join(str(tmp_path), "scenario.py")
    with open(script, "w") as f:
        f.write(PRELUDE + textwrap.dedent(body) + "\nprint('RESULT ' + json.dumps(out))\n")
    env = dict(os.environ, SAMGRAPH_LOG_LEVEL="warn", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, script, cfg_path], capture_output=True, text=True, timeout=300, env=env)
    res = None
    for line in r.stdout.splitlines():
        if line.startswith("RESULT "):
            res = json.loads(line[7:])
    return r, res


def test_data_init_loads_the_reference_dataset_format(tmp_path, dataset):
    r, out = run(tmp_path, config(dataset["path"]), """
        sam.config(cfg)
        sam.data_init()
        out.update(num_epoch=sam.num_epoch(), steps=sam.steps_per_epoch(), num_class=sam.num_class(),
                   feat_dim=sam.feat_dim())
        feat, label = sam.get_dataset_feat(), sam.get_dataset_label()
        out.update(feat_shape=list(feat.shape), label_shape=list(label.shape), feat_dtype=str(feat.dtype),
                   label_dtype=str(label.dtype), feat_sum=float(feat.double().sum()), label_sum=int(label.sum()),
                   feat_is_cpu=feat.device.type == "cpu")
        sam.shutdown()
    """)
    assert r.returncode == 0, r.stderr[-2000:]
    n_train = len(dataset["train_set"])
    assert out["num_epoch"] == 3 and out["steps"] == (n_train + 127) // 128        # engine.h:49-53
    assert out["num_class"] == dataset["num_class"] and out["feat_dim"] == dataset["feat_dim"]
    assert out["feat_shape"] == list(dataset["feat"].shape) and out["label_shape"] == [len(dataset["label"])]
    assert out["feat_dtype"] == "torch.float32" and out["label_dtype"] == "torch.int64" and out["feat_is_cpu"]
    assert out["feat_sum"] == float(dataset["feat"].astype(np.float64).sum())
    assert out["label_sum"] == int(dataset["label"].sum())


@pytest.mark.parametrize("policy", ["degree_hop", "fake_optimal"])
def test_data_init_builds_the_offline_tool_rankings_when_their_file_is_absent(tmp_path, dataset, oracle, policy):
    """engine.cc:233-244 only loads cache_by_degree_hop.bin / cache_by_fake_optimal.bin; without the file the loader
    builds the ranking with the host restatement of the reference's offline tool (include/fgnn_dataset_tools.h),
    and with the file it uses the file."""
    body = """
        import ctypes
        sam.config(cfg)
        sam.data_init()
        lib = ctypes.CDLL(sam.c_lib.__file__)
        lib.fgnn_rt_dataset_ranking.restype = ctypes.POINTER(ctypes.c_uint32)
        n = ctypes.c_size_t(0)
        p = lib.fgnn_rt_dataset_ranking(ctypes.byref(n))
        out["rank"] = np.ctypeslib.as_array(p, shape=(n.value,)).tolist() if p else None
        sam.shutdown()
    """
    import samgraph.common as sc
    fname = os.path.join(dataset["path"], "cache_by_%s.bin" % policy)
    assert not os.path.exists(fname)
    cfg = config(dataset["path"], cache_policy=policy, _cache_policy=sc.cache_policies[policy])
    r, out = run(tmp_path, cfg, body)
    assert r.returncode == 0, r.stderr[-2000:]
    if policy == "degree_hop":
        want = oracle.rank_by_degree_hop(dataset["indptr"], dataset["indices"], dataset["train_set"])
    else:
        want, _ = oracle.rank_by_fake_optimal(dataset["indptr"], dataset["indices"], dataset["train_set"], order_threads=48)
    assert out["rank"] == want.tolist()
    try:                                            # a file, when present, wins
        np.arange(len(want), dtype=np.uint32)[::-1].tofile(fname)
        r, out = run(tmp_path, cfg, body)
        assert r.returncode == 0, r.stderr[-2000:]
        assert out["rank"] == list(range(len(want) - 1, -1, -1))
    finally:
        os.remove(fname)


def test_profiler_log_api_round_trips(tmp_path, dataset):
    r, out = run(tmp_path, config(dataset["path"]), """
        sam.config(cfg)
        sam.data_init()
        sam.log_step(0, 1, sam.kLogL1TrainTime, 2.5)
        sam.log_step_add(0, 1, sam.kLogL1TrainTime, 1.0)
        sam.log_step(2, 5, sam.kLogL1ConvertTime, 0.25)
        sam.log_epoch_add(1, sam.kLogEpochTrainTime, 4.0)
        sam.log_epoch_add(1, sam.kLogEpochTrainTime, 0.5)
        out.update(step=sam.get_log_step_value(0, 1, sam.kLogL1TrainTime),
                   step2=sam.get_log_step_value(2, 5, sam.kLogL1ConvertTime),
                   other=sam.get_log_step_value(0, 2, sam.kLogL1TrainTime),
                   epoch=sam.get_log_epoch_value(1, sam.kLogEpochTrainTime),
                   epoch_other=sam.get_log_epoch_value(0, sam.kLogEpochTrainTime))
        sam.report_step(0, 1); sam.report_step_average(0, 1); sam.report_epoch(1); sam.report_epoch_average(1)
        sam.trace_step_begin_now(1, sam.kL1Event_Train); sam.trace_step_end_now(1, sam.kL1Event_Train)
        sam.shutdown()
    """)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out == {"step": 3.5, "step2": 0.25, "other": 0.0, "epoch": 4.5, "epoch_other": 0.0}   # profiler.cc LogStep/Add


@pytest.mark.parametrize("fanout,batch,expect_steps", [([25, 10], 700, 1), ([5, 10, 15], 699, 2), ([3], 1, 700)])
def test_step_arithmetic_and_list_values(tmp_path, dataset, fanout, batch, expect_steps):
    cfg = config(dataset["path"], fanout=fanout, num_fanout=len(fanout), num_layer=len(fanout), batch_size=batch,
                 num_epoch=2)
    r, out = run(tmp_path, cfg, """
        sam.config(cfg)
        sam.data_init()
        out.update(steps=sam.steps_per_epoch(), num_epoch=sam.num_epoch())
        sam.shutdown()
    """)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out == {"steps": expect_steps, "num_epoch": 2}


@pytest.mark.parametrize("case,needle", [
    ("missing_key", "batch_size"), ("missing_dataset", "meta.txt"), ("cpu_sampler", "CPU"),
    ("switch_init", "switch_init"), ("short_file", "smaller")])
def test_violated_checks_abort_like_the_reference(tmp_path, dataset, case, needle):
    """No error codes on this ABI: a violated CHECK logs and abort()s the process (logging.h:32-62,
    logging.cc:69-73); the multi-process parent learns about it from the child's exit status."""
    cfg = config(dataset["path"])
    body = "sam.config(cfg)\nsam.data_init()\n"
    if case == "missing_key":
        del cfg["batch_size"]
    elif case == "missing_dataset":
        cfg["dataset_path"] = os.path.join(str(tmp_path), "nowhere")
    elif case == "cpu_sampler":
        cfg.update(_arch=0, arch="arch0", sampler_ctx="cpu:0", trainer_ctx="cuda:0")
        body = "sam.config(cfg)\nsam.init()\n"
    elif case == "switch_init":
        body += "sam.switch_init(0, 'cuda:0', 0.1)\n"
    elif case == "short_file":
        import shutil
        bad = os.path.join(str(tmp_path), "short")
        shutil.copytree(dataset["path"], bad)
        with open(os.path.join(bad, "indices.bin"), "r+b") as f:
            f.truncate(1000)
        cfg["dataset_path"] = bad
    r, out = run(tmp_path, cfg, body)
    assert r.returncode != 0 and out is None
    assert r.returncode in (-6, 134), "expected SIGABRT, got %d\n%s" % (r.returncode, r.stderr[-1500:])
    assert needle.lower() in (r.stderr + r.stdout).lower(), r.stderr[-1500:]


def test_tensor_getters_without_a_batch_raise_instead_of_aborting(tmp_path, dataset):
    """The DLPack getters of samgraph.torch.c_lib (adapter.cc:48-192 equivalents): before any get_next_batch there
    is no current batch — a Python exception, the process stays alive."""
    r, out = run(tmp_path, config(dataset["path"]), """
        sam.config(cfg)
        sam.data_init()
        errs = {}
        for name, args in (("get_graph_feat", (0,)), ("get_graph_label", (0,)), ("get_graph_row", (0, 0)),
                           ("get_graph_col", (0, 1)), ("get_graph_data", (0, 0)), ("get_graph_csc", (0, 0)),
                           ("get_graph_input_nodes", (0,)), ("get_graph_output_nodes", (0,))):
            try:
                getattr(sam, name)(*args)
                errs[name] = "no error"
            except RuntimeError as e:
                errs[name] = "RuntimeError: " + str(e)
        out.update(errs)
        out["alive"] = True
        sam.shutdown()
    """)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.pop("alive") is True
    assert len(out) == 8
    for name, msg in out.items():
        assert msg.startswith("RuntimeError: samgraph: no current batch"), (name, msg)
