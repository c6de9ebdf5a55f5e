"""bench.py's contract on the CPU: the reference arm prints one JSON line with
the keys the driver reads, runs nothing of the product (the CUDA library is
never mapped), and the native arm refuses to run without a GPU instead of
falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          env=dict(os.environ, **(env or {})), timeout=600)


def test_reference_arm_line_and_no_product_code():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], env={"LD_DEBUG": "files"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "likelihood_evals_per_s" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C4-1024: sie_plus_shear+sersic+sersic+sky, 1024x1024")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the dynamic loader's file trace (LD_DEBUG=files goes to stderr): oracle libraries yes, the CUDA library never
    assert "liblensed_cuda" not in r.stderr
    assert "oracle" in r.stderr


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
