"""Developer timing script (GPU box): the stem kernel alone (fp32 NCHW input and decoded uint8 HWC input), N images."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = 112
lib = _lib.load()
P = _lib.ptr
x = torch.randn(n, 3, S, S, device="cuda").clamp_(-1, 1)
u = torch.randint(0, 256, (n, S, S, 3), dtype=torch.uint8, device="cuda")
w = torch.randn(27, 64, device="cuda") * 0.1
b = torch.zeros(64, device="cuda")
a = torch.full((64,), 0.25, device="cuda")
out = torch.empty(n * (S + 1) * (S + 1), 64, dtype=torch.bfloat16, device="cuda")
res = {}
for name, fn in (("fp32_nchw", lambda: lib.ffr_stem_fwd(P(x), P(w), P(b), P(a), P(out), n, S, _lib.stream_ptr())),
                 ("u8_hwc", lambda: lib.ffr_stem_u8_fwd(P(u), None, 1, P(w), P(b), P(a), P(out), n, S, _lib.stream_ptr()))):
    for _ in range(3):
        _lib.check(fn())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        _lib.check(fn())
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    gb = (out.numel() * 2 + (x.numel() * 4 if name == "fp32_nchw" else u.numel())) / 1e9
    res[name] = {"us": us, "GBps": gb / (us * 1e-6)}
    print(name, json.dumps(res[name]))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/stem_bench_%d.json" % n, "w"))
