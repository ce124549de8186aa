"""bench.py's CPU-side contract: the reference arm prints one JSON line with the agreed keys, the scene generator is
deterministic, and the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line(ref):
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--fills", "300", "--width", "640", "--height", "360")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "unit", "value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "Mpix/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0
    assert "workload" in d["config"] and d["config"]["fills_per_step"] == 300      # the whole step, not a sample


def test_scene_generator_is_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    a, keep_a = bench.make_config1_scene(500, 3840, 2160, seed=1234)
    b, keep_b = bench.make_config1_scene(500, 3840, 2160, seed=1234)
    va = np.ctypeslib.as_array(a.vertices, (a.vertex_count * 2,))
    vb = np.ctypeslib.as_array(b.vertices, (b.vertex_count * 2,))
    assert a.fill_count == b.fill_count == 500 and np.array_equal(va, vb)
    kinds = {a.fills[i].geom for i in range(500)}
    styles = {a.fills[i].style for i in range(500)}
    rules = {a.fills[i].fill_rule for i in range(500)}
    assert kinds == {2, 3} and styles == {1, 2, 3} and rules == {0, 1}


def test_gpu_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
