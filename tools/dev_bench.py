"""Developer timing script (GPU box): backbone forward time, total and per library call. Not the contract bench."""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone as ob  # weights generator only
from ffr_net_b200.backbone import Backbone


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    m = Backbone(50, 0.6, "ir_se")
    m.load_state_dict(ob.synth_backbone_state_dict(0))
    m = m.cuda().eval()
    x = ob.synth_faces(min(n, 64), 0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        iters = 10
        for _ in range(iters):
            m(x)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / iters
        print("N=%d backbone fwd %.3f ms -> %.0f img/s" % (n, ms, n / ms * 1e3))
        # per-call breakdown
        m._profile = []
        e0 = torch.cuda.Event(enable_timing=True)
        e0.record()
        m(x)
        torch.cuda.synchronize()
        prev = e0
        agg = collections.OrderedDict()
        rows = []
        for what, ev in m._profile:
            dt = prev.elapsed_time(ev)
            rows.append((what, dt))
            agg[what] = agg.get(what, 0.0) + dt
            prev = ev
        m._profile = None
        print(json.dumps(agg))
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/dev_bench_%d.json" % n, "w") as f:
            json.dump({"n": n, "ms": ms, "img_s": n / ms * 1e3, "agg": agg, "rows": rows}, f)


if __name__ == "__main__":
    main()
