"""Developer timing script (GPU box): backbone forward time, total and per library call. Not the contract bench."""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
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
        out = {"n": n, "ms": ms, "img_s": n / ms * 1e3, "agg": agg, "rows": rows}
        # ---- RecNet stage ----
        try:
            from ffr_net_b200 import synth as orr
            from ffr_net_b200.recnet import RecNet
            rec = RecNet()
            rec.load_state_dict(orr.synth_recnet_state_dict(0))
            rec = rec.cuda().eval()
            for _ in range(3):
                rec.embed_from_images(m, x)
            torch.cuda.synchronize()
            s.record()
            for _ in range(iters):
                rec.embed_from_images(m, x)
            e.record()
            torch.cuda.synchronize()
            ms2 = s.elapsed_time(e) / iters
            print("N=%d backbone+RecNet fwd %.3f ms -> %.0f img/s" % (n, ms2, n / ms2 * 1e3))
            y, _ = m(x)
            rec._profile = []
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rec(y)
            torch.cuda.synchronize()
            prev = e0
            agg2 = collections.OrderedDict()
            for what, ev in rec._profile:
                agg2[what] = agg2.get(what, 0.0) + prev.elapsed_time(ev)
                prev = ev
            rec._profile = None
            print(json.dumps(agg2))
            out.update({"ms_with_recnet": ms2, "img_s_with_recnet": n / ms2 * 1e3, "recnet_agg": agg2})
        except Exception as ex:  # keep the backbone numbers even if the RecNet stage fails
            print("recnet stage failed:", repr(ex))
        with open("gpurun_out/dev_bench_%d.json" % n, "w") as f:
            json.dump(out, f)


if __name__ == "__main__":
    main()
