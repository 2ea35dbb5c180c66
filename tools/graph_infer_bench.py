"""Developer timing script (GPU box): eval step (backbone + RecNet, N images) eager vs replayed from a CUDA graph."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
from ffr_net_b200 import synth as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(ob.synth_backbone_state_dict(0))
enc = enc.cuda().eval()
rec = RecNet()
rec.load_state_dict(orr.synth_recnet_state_dict(0))
rec = rec.cuda().eval()
x = ob.synth_faces(min(n, 64), 0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


with torch.no_grad():
    ms_eager = timed(lambda: rec.embed_from_images(enc, x))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            rec.embed_from_images(enc, x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = rec.embed_from_images(enc, x)
    ms_graph = timed(g.replay)
    ref = rec.embed_from_images(enc, x)
    g.replay()
    torch.cuda.synchronize()
    print("N=%d eager %.3f ms (%.0f img/s) | graph replay %.3f ms (%.0f img/s) | max diff %.2e" %
          (n, ms_eager, n / ms_eager * 1e3, ms_graph, n / ms_graph * 1e3, (out - ref).abs().max().item()))
