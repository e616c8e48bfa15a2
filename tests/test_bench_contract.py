"""bench.py contract checks that run without a GPU: the reference arm's JSON line, the loud failure of the
product arm without CUDA, and the contract flop figures of SURVEY.md §8d."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_contract_flops_per_iteration():
    b = _bench()
    # F_it(256, 60, 20, 16) = 9.26 MFLOP, F_it(256, 64, 1, 16) = 0.49 MFLOP
    assert abs(b.stage_flops(256, 60, 20, 16, False) - 9.26e6) < 0.02e6
    assert abs(b.stage_flops(256, 64, 1, 16, False) - 0.49e6) < 0.01e6
    # the explicit-inverse branch wins once m is large (min of the two formulations)
    assert b.stage_flops(256, 1024, 20, 16, False) == 8 * 256 * 256 * 20 + 24 * 256 * 1024 * 20 + 16 * 16 * 256 * 20
    assert set(b.WORKLOADS) == {"config0", "config1", "config3", "config4", "config5"}


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config0",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1 and line["scaling"] == "weak"
    assert "inferLowRankV4_multi" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout) and "no CPU fallback" in (out.stderr + out.stdout)
