"""CPU tests of the bench.py contract that need no GPU: the reference arm prints one well-formed JSON line, and the
product arm refuses to run (loudly, non-zero exit) when there is no CUDA device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra=None, timeout=600):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert d["metric"] == "IR-SE50+RecBlock embeddings/s (bs512)" and d["value"] > 0
    # "reference": the unmodified reference modules staged under baseline/_ref by __graft_entry__.build(); "port": the oracle
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "models", "recnet.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--gpus", "2"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--no-train"])
    assert r.returncode != 0
    assert "cuda" in (r.stderr + r.stdout).lower()
    assert not any(l.startswith("{") for l in r.stdout.splitlines())      # no benchmark line from a CPU fallback
