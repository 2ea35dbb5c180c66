"""Developer timing script (GPU box): throughput of the chunked multi-stream forward (ffr_net_b200/streams.py) for
FFR_STREAMS = 1..4 at one batch size. Not the contract bench."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
from ffr_net_b200 import synth as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    m = Backbone(50, 0.6, "ir_se")
    m.load_state_dict(ob.synth_backbone_state_dict(0))
    m = m.cuda().eval()
    rec = RecNet()
    rec.load_state_dict(orr.synth_recnet_state_dict(0))
    rec = rec.cuda().eval()
    x = ob.synth_faces(min(n, 64), 0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()
    out = {"n": n, "runs": []}
    with torch.no_grad():
        for k in (1, 2, 3, 4, 1, 2):
            os.environ["FFR_STREAMS"] = str(k)
            ms_b = timed(lambda: m(x))
            ms_e = timed(lambda: rec.embed_from_images(m, x))
            row = {"streams": k, "backbone_ms": ms_b, "backbone_img_s": n / ms_b * 1e3,
                   "embed_ms": ms_e, "embed_img_s": n / ms_e * 1e3}
            print(json.dumps(row))
            out["runs"].append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/streams_sweep_%d.json" % n, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
