"""Data-parallel training step on >= 2 GPUs over NCCL (run under `gpurun --gpus 2`): the averaged gradients equal the
single-GPU average of the per-shard gradients (BatchNorm statistics per rank, models/trainer.py:70,72 data_parallel
semantics), the bucketed / overlapped exchange equals the single flat all-reduce bit for bit, and a CUDA-graph-captured
DP step leaves every rank with identical parameters. Skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_data_parallel_step_two_ranks(lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DP_RESULT ")]
    assert r.returncode == 0 and lines, (r.stdout[-2000:], r.stderr[-3000:])
    out = json.loads(lines[-1][len("DP_RESULT "):])
    print(out)
    assert out["overlap_equals_flat"]
    assert out["worst_rel_err_vs_single_gpu_average"] <= 1e-6
    assert out["params_identical_across_ranks"] and out["finite"]
