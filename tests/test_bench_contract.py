"""bench.py's driver-facing contract that can be checked without a GPU: the reference arm prints one JSON
line with the agreed keys, and the product arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ)
    e.pop("RANK", None); e.pop("WORLD_SIZE", None); e.pop("LOCAL_RANK", None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=e)


def test_reference_arm_json_line():
    r = run("--impl", "reference", "--queries", "2", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout  # exactly one line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "UIDs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert cb["cores"] == min(2, os.cpu_count())  # the threads that are actually busy: one per query
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_same_config_at_n_gpus():
    """At N GPUs the reference arm runs the whole job's N x Q queries on rank 0 and reports the GPU arm's config."""
    r = run("--impl", "reference", "--gpus", "2", "--queries", "2", "--steps", "1", "--warmup", "1",
            env={"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["n_gpus"] == 2 and d["config"]["queries_per_gpu"] == 2
    assert "2 GPU(s)" in d["config"]["parallelism"]
    assert d["cpu_baseline"]["cores"] == min(4, os.cpu_count())
    assert "4 queries (2 x 2)" in d["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--gpus", "2", "--queries", "2", "--steps", "1", "--warmup", "0",
            env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = run("--steps", "1", "--warmup", "0", "--queries", "1")
    assert r.returncode != 0
    assert "CUDA" in (r.stderr + r.stdout)
