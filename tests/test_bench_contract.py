"""bench.py's CPU-runnable leg: `--impl reference` must print one JSON line with the contract's keys (the GPU arm needs a
device and is exercised by the driver)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")     # torchrun exports this to its workers: the arm must not obey it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--width", "96", "--height", "96", "--spp", "4", "--cpu-seconds", "0.5"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and "sample" in cb
    ncpu = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncpu                      # all host cores, whatever OMP_NUM_THREADS says
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_the_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback anywhere on the product path: without a CUDA device the b200 arm must stop with an error, not print a line."""
    from conftest import HAS_GPU
    if HAS_GPU:
        import pytest
        pytest.skip("needs a machine WITHOUT a GPU")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline",
                        "--no-per-config"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
