"""Wall-clock (%globaltimer) entry / exit of consecutive launches of the pair conv kernel: in-kernel span and the gap
between one launch's last CTA exit and the next launch's first CTA entry, with programmatic dependent launch on / off."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib, packing

lib = _lib.load()
P = _lib.ptr
N, S, cin, cout = 512, 14, 256, 256
g = torch.Generator(device="cuda").manual_seed(0)
rows = N * (S + 1) * (S + 1)
x = (torch.randn(rows, cin, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
wp = packing.pack_conv(w)
bias9 = torch.zeros(9, cout, device="cuda")
slope = torch.full((cout,), 0.25, device="cuda")
out = torch.empty(rows, cout, dtype=torch.bfloat16, device="cuda")
st = _lib.stream_ptr()
res = {}
for pdl in (3, 0):
    lib.ffr_debug_set_pdl(pdl)
    K = 8
    bufs = []
    for i in range(K):
        b = torch.zeros(16, dtype=torch.int64, device="cuda")
        b[11] = b[12] = 2 ** 62
        bufs.append(b)
    for _ in range(3):
        _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        lib.ffr_debug_set_counters(P(bufs[i]))
        _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(x), N, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, st))
    e1.record()
    lib.ffr_debug_set_counters(None)
    torch.cuda.synchronize()
    d = [b.tolist() for b in bufs]
    t0 = d[0][11]
    rowsr = [dict(entry=r[11] - t0, mma_begin=r[12] - t0, mma_end=r[13] - t0, exit=r[14] - t0) for r in d]
    gaps = [rowsr[i + 1]["entry"] - rowsr[i]["exit"] for i in range(K - 1)]
    spans = [r["exit"] - r["entry"] for r in rowsr]
    res["pdl=%d" % pdl] = dict(event_ms_per_launch=e0.elapsed_time(e1) / K, spans_ns=spans, gaps_ns=gaps, launches=rowsr)
    print("pdl", pdl, "event us/launch %.2f" % (e0.elapsed_time(e1) / K * 1e3), "spans", spans, "gaps", gaps)
lib.ffr_debug_set_pdl(-1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/launch_gaps.json", "w"), indent=1)
