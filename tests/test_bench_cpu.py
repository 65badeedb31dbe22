"""bench.py contract pieces that run without a GPU: the reference arm (the reference's own CPU sampler + extractor,
oracle/_ref) prints the contract's JSON line with the same metric / unit / config.workload as our arm."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(ref):
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ci-64k",
                        "--steps", "3", "--warmup", "1", "--empty-feat", "14"], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [x for x in r.stdout.splitlines() if x.startswith("{")][-1]
    out = json.loads(line)
    sys.path.insert(0, ROOT)
    import bench
    assert out["impl"] == "reference" and out["metric"] == bench.METRIC and out["unit"] == bench.UNIT
    assert out["higher_is_better"] is True and out["n_gpus"] == 1 and out["value"] > 0
    assert out["e2e"] == {"value": out["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = out["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == out["value"] and cb["cores"] >= 1 and cb["sample"]

    class A:
        workload, cache_pct = "ci-64k", 1.0
    assert out["config"]["workload"] == bench.workload_name(A)      # same workload name as our arm's line


@pytest.mark.parametrize("sample_type,fanout", [("random_walk", None), ("weighted_khop", "10,5"), ("khop0", "5,10,15")])
def test_reference_arm_other_samplers(ref, sample_type, fanout):
    """Configs #3 / #5 (PinSAGE random walk, weighted k-hop): the reference has no CPU code for them (empty stubs,
    SURVEY §8c), so the arm times the oracle port and says so (kind "port"); khop0 with another fanout list is the
    reference's own code again."""
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "ci-64k", "--steps", "2",
           "--warmup", "1", "--empty-feat", "14", "--sample-type", sample_type]
    if fanout:
        cmd += ["--fanout", fanout]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="4"))
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([x for x in r.stdout.splitlines() if x.startswith("{")][-1])
    cb = out["cpu_baseline"]
    assert out["impl"] == "reference" and out["value"] > 0 and cb["value"] == out["value"] and cb["steps"] == 2
    assert cb["kind"] == ("reference" if sample_type == "khop0" else "port")
    assert sample_type in cb["sample"] and out["config"]["sample_type"] == sample_type


def test_rank_nonzero_of_reference_arm_is_silent():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
