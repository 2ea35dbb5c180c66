"""Developer experiment (GPU box): backbone 3x3/s1 convolutions with pixel-major tiles on the small maps
(ffr_debug_set_pixmajor_backbone: S <= 0 / 7 / 14 / 28) — time and agreement with the row-major sliding-window path."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
from ffr_net_b200 import _lib
from ffr_net_b200.backbone import Backbone

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
lib = _lib.load()
m = Backbone(50, 0.6, "ir_se")
m.load_state_dict(ob.synth_backbone_state_dict(0))
m = m.cuda().eval()
x = ob.synth_faces(min(n, 64), 0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()
rows = []
ref = None
with torch.no_grad():
    for max_s in (0, 7, 14, 28, 0, 14):
        lib.ffr_debug_set_pixmajor_backbone(max_s)
        for _ in range(3):
            y, f = m(x)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            y, f = m(x)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        if ref is None:
            ref = (y.clone(), f.clone())
        row = {"pix_max_s": max_s, "ms": ms, "img_s": n / ms * 1e3,
               "y_rel": ((y - ref[0]).abs().max() / ref[0].abs().max()).item(), "f_abs": (f - ref[1]).abs().max().item()}
        print(json.dumps(row))
        rows.append(row)
lib.ffr_debug_set_pixmajor_backbone(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/pix_backbone_bench_%d.json" % n, "w"), indent=1)
