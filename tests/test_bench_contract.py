"""bench.py's JSON contract on the arm that runs without a GPU: `--impl reference` (the CPU restatement timed on the
host cores) must print ONE line with the keys the driver reads; the GPU arm refuses to run without a device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--cells", "642", "--levels", "10", "--steps", "2", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "cell-columns/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f64" and "workload" in d["config"] and d["n_gpus"] == 1 and d["steps"] == 2
    cb = d["cpu_baseline"]
    # "reference" where oracle/_ref (the reference's own source, transliterated and compiled) is built, else the port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] - 642 * 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]


def test_reference_arm_decomposed(tmp_path):
    """N > 1: rank 0 steps all N blocks of the decomposition in lock step (nothing extrapolated) with the C++ restatement."""
    r = _run("--impl", "reference", "--gpus", "2", "--cells", "642", "--levels", "10", "--steps", "2", "--warmup", "3",
             env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "MPASB_CACHE": str(tmp_path)})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    cb = d["cpu_baseline"]
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and cb["kind"] == "port" and cb["extrapolated"] is False
    assert "all 2 blocks" in cb["sample"] and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--cells", "642", "--levels", "10", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--cells", "642", "--levels", "10", "--steps", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
