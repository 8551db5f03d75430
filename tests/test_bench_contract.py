"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints exactly ONE JSON line with
the agreed keys (it times the reference's CPU path, so it runs anywhere), and our arm refuses to run without a B200
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, stdout=subprocess.PIPE,
                          stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "3", "--streams", "8", "--seconds", "0.25")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "samples/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "3", "--streams", "2", "--seconds", "0.1")
    assert r.returncode != 0
    assert r.stdout.strip() == ""          # no number without the CUDA path
    assert "no CUDA device" in r.stderr or "no CPU fallback" in r.stderr
