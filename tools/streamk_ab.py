"""A/B of the stream-K schedule of the CTA-pair window kernel (ffr_set_conv_scratch) on the backbone's 256-wide layers at
batch 512: launch time (CUDA events, back-to-back launches) and the per-role wait counters, scratch off / on interleaved."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib, packing

lib = _lib.load()
P = _lib.ptr
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
names = ["mma_wait_tmem", "mma_wait_a", "mma_wait_b", "mma_total", "tma_wait_a", "tma_wait_b", "tma_total", "-",
         "epi_wait", "epi_total", "ctas"]
scratch = torch.zeros(lib.ffr_conv_scratch_bytes(), dtype=torch.uint8, device="cuda")
lib.ffr_debug_set_streamk(1)      # opt-in schedule; without a registered scratch the launches run whole items
res = []
for (S, cin, cout) in [(14, 256, 256), (14, 256, 512), (7, 512, 512), (28, 128, 256)]:
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = N * (S + 1) * (S + 1)
    x = (torch.randn(rows, cin, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
    wp = packing.pack_conv(w)
    bias9 = torch.zeros(9, cout, device="cuda")
    slope = torch.full((cout,), 0.25, device="cuda")
    out = torch.empty(rows, cout, dtype=torch.bfloat16, device="cuda")
    st = _lib.stream_ptr()

    def call():
        _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))

    for rep in range(2):
        for sk in (0, 1):
            lib.ffr_set_conv_scratch(P(scratch) if sk else None, scratch.numel() if sk else 0)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                call()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            used = lib.ffr_debug_last_streamk()
            dbg = torch.zeros(16, dtype=torch.int64, device="cuda")
            dbg[11] = dbg[12] = 2 ** 62
            lib.ffr_debug_set_counters(P(dbg))
            call()
            torch.cuda.synchronize()
            lib.ffr_debug_set_counters(None)
            d = dbg.tolist()
            ctas = max(1, d[10])
            flop = 2.0 * N * S * S * cout * cin * 9
            row = dict(S=S, cin=cin, cout=cout, scratch=sk, streamk_used=used, ms=ms, tflops=flop / ms / 1e9,
                       **{n: d[i] / ctas for i, n in enumerate(names) if n != "-"})
            row.update(t_mma_begin_ns=d[12] - d[11], t_mma_end_ns=d[13] - d[11], t_exit_ns=d[14] - d[11], mma_total_max_cycles=d[15])
            res.append(row)
            print(json.dumps(row))
lib.ffr_set_conv_scratch(None, 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/streamk_ab.json", "w"), indent=1)
