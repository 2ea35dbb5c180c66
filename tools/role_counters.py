"""Per-role barrier-wait breakdown of the sliding-window conv kernel for the backbone's conv shapes (GPU box)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib, packing, layout

lib = _lib.load()
P = _lib.ptr
N = 512
res = []
names = ["mma_wait_tmem", "mma_wait_a", "mma_wait_b", "mma_total", "tma_wait_a", "tma_wait_b", "tma_total", "-",
         "epi_wait", "epi_total", "ctas"]
SHAPES = [(112, 64, 64), (56, 64, 64), (56, 64, 128), (28, 128, 128), (28, 128, 256), (14, 256, 256), (14, 256, 512),
          (7, 512, 512)]
for (S, cin, cout, pair) in [(S, ci, co, pr) for (S, ci, co) in SHAPES for pr in ((0, 1) if co >= 256 else (0,))]:
    lib.ffr_debug_set_pair(pair)
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = N * (S + 1) * (S + 1)
    x = (torch.randn(rows, cin, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
    wp = packing.pack_conv(w)
    bias9 = torch.zeros(9, cout, device="cuda")
    slope = torch.full((cout,), 0.25, device="cuda")
    out = torch.empty(rows, cout, dtype=torch.bfloat16, device="cuda")
    dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
    dbg[11] = dbg[12] = 2 ** 62
    st = _lib.stream_ptr()
    for _ in range(3):
        _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    lib.ffr_debug_set_counters(P(dbg))
    _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))
    torch.cuda.synchronize()
    lib.ffr_debug_set_counters(None)
    d = dbg.tolist()
    ctas = max(1, d[10])
    flop = 2.0 * N * S * S * cout * cin * 9
    row = dict(S=S, cin=cin, cout=cout, pair=pair, ms=ms, tflops=flop / ms / 1e9, **{n: d[i] / ctas for i, n in enumerate(names) if n != "-"})
    if pair and d[14] > 0:      # wall-clock phases of the pair kernel (ns since the earliest CTA entry)
        row.update(t_mma_begin_ns=d[12] - d[11], t_mma_end_ns=d[13] - d[11], t_exit_ns=d[14] - d[11], mma_total_max_cycles=d[15])
    res.append(row)
    print(json.dumps(row))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/role_counters.json", "w"), indent=1)
