"""CPU: the measurement contract of bench.py that can be checked without a GPU — the reference arm prints exactly one
JSON line with the agreed keys (timing the compiled reference on the host cores), and our arm refuses to run without a
device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

import refdump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


@pytest.mark.skipif(not refdump.have_ref(), reason="compiled reference (oracle/_ref) not built")
def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--config", "cfg1", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cholesky_factor_gflops" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_on_other_ranks_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--config", "cfg1", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, BENCH, "--config", "cfg1", "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert r.stdout.strip() == ""              # nothing that could be mistaken for a result
