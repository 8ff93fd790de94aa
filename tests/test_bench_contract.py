"""CPU checks of bench.py's contract with the driver: the reference arm (the one leg that runs without a GPU)
prints ONE JSON line with the keys and meanings the contract names, for the headline workload and for the
occlusion-solver workload; under a multi-rank launch only rank 0 works and prints; without a GPU the product
arm refuses to run instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_line_small_shape():
    """--impl reference on a small frame shape (same code path as 1080p, seconds instead of minutes)."""
    p = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--nx", "320", "--ny", "240"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = json_lines(p.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "frame-pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"


def test_reference_arm_only_rank_zero_prints():
    p = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--nx", "160", "--ny", "120", "--gpus", "2"],
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and json_lines(p.stdout) == []


def test_product_arm_refuses_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for extra in ([], ["--workload", "occ"], ["--workload", "band4k"]):
        p = run(["--steps", "1", "--warmup", "0"] + extra)
        assert p.returncode != 0
        assert "no CUDA device" in (p.stderr + p.stdout) or "no CPU fallback" in (p.stderr + p.stdout)
        assert json_lines(p.stdout) == []
