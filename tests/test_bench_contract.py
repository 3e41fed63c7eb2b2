"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the reference
path) prints ONE JSON line with the agreed keys, and the product arm refuses to run without CUDA
(there is no CPU fallback to time by accident)."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args, timeout=300):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=str(ROOT))


def test_reference_arm_prints_one_contract_line():
    r = _run("--impl", "reference", "--workload", "stories15m", "--steps", "8", "--warmup", "3")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_tokens_per_s" and d["unit"] == "tokens/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 8 and d["warmup"] == 3
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "stories15M" in cb["sample"]
    assert d["config"]["workload"].startswith("stories15M") and d["gpu_launches"] == 0
    assert d["vs_baseline"] is None  # BASELINE.md publishes a number for TinyLlama-1.1B fp32 only


def test_reference_arm_other_ranks_exit_quietly(monkeypatch):
    monkeypatch.setenv("RANK", "1")
    monkeypatch.setenv("WORLD_SIZE", "2")
    r = _run("--impl", "reference", "--gpus", "2", "--workload", "stories15m", "--steps", "4", "--warmup", "3")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--workload", "stories15m", "--steps", "4", "--warmup", "3", "--no-cpu-baseline")
    assert r.returncode != 0 and r.stdout.strip() == ""
