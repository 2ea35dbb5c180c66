"""tcgen05.mma issue-rate micro-benchmark driver (GPU box). Writes gpurun_out/mma_bench.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib

lib = _lib.load()
res = []
iters = 4096
for grid in (1, 148):
    for M in (128, 64):
        for N in (32, 64, 128, 256):
            for n_acc in (1, 2, 4):
                if n_acc * N > 512:
                    continue
                out = torch.zeros(grid, dtype=torch.int64, device="cuda")
                _lib.check(_lib.load_probe().ffr_debug_mma_bench(_lib.ptr(out), M, N, n_acc, iters, grid, _lib.stream_ptr()))
                torch.cuda.synchronize()
                cyc = out.float().mean().item() / iters
                res.append(dict(grid=grid, M=M, N=N, n_acc=n_acc, cycles_per_mma=cyc))
                print(res[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/mma_bench.json", "w"), indent=1)
