"""Developer timing script (GPU box): the K=64 GEMM that produces M_channel (n*512 x 512 outputs, bias + sigmoid)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib

lib = _lib.load()
P = _lib.ptr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
h5 = (torch.randn(n * 512, 64, device="cuda") * 0.5).bfloat16()
w8 = (torch.randn(512, 64, device="cuda") * 0.1).bfloat16()
b8 = torch.zeros(512, device="cuda")
out = torch.empty(n * 512, 512, dtype=torch.bfloat16, device="cuda")
names = ["mma_wait_tmem", "mma_wait_a", "mma_wait_b", "mma_total", "tma_wait_a", "tma_wait_b", "tma_total", "-",
         "epi_wait", "epi_total", "ctas"]
rows = []
for label, flags in (("bias+sigmoid", 0x1 | 0x80), ("bias", 0x1), ("plain", 0x0)):
    def run():
        _lib.check(lib.ffr_conv_gemm(P(h5), n * 512, 64, 64, P(w8), 64, 512, 1, None, None, n * 512, 64, 1, 1, 0, n, flags,
                                     P(b8), None, P(out), 512, 0, None, None, None, 0, None, 1, None, 0, 0, 0,
                                     _lib.stream_ptr()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    lib.ffr_debug_set_counters(P(dbg))
    run()
    torch.cuda.synchronize()
    lib.ffr_debug_set_counters(None)
    d = dbg.tolist()
    ctas = max(1, d[10])
    row = dict(case=label, us=us, gbps_out=n * 512 * 512 * 2 / us / 1e3, **{k: d[i] / ctas for i, k in enumerate(names) if k != "-"})
    print(json.dumps(row))
    rows.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/mchannel_bench.json", "w"), indent=1)
