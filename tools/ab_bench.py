"""A/B timing of the eval step (IR-SE50 + RecNet, batch 512) under the launch-mode switches, interleaved in one process
on one box (box-to-box clocks differ by a few %): programmatic dependent launch on/off x CTA pairs on/off."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import _lib, synth
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet

lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(synth.synth_backbone_state_dict(0))
enc = enc.cuda().eval()
rec = RecNet()
rec.load_state_dict(synth.synth_recnet_state_dict(0))
rec = rec.cuda().eval()
x = synth.synth_faces(64, seed=0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()


def run(iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        s.record()
        for _ in range(iters):
            rec.embed_from_images(enc, x)
        e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


res = {}
with torch.no_grad():
    for _ in range(5):
        rec.embed_from_images(enc, x)
# (pdl, pair, fused SE, lean epilogue, stream-K, strip stem, RecNet branches on two streams)
CONFIGS = [(-1, -1, True, 1, 0, 1, 1), (-1, -1, True, 1, 0, 1, 0), (-1, -1, True, 1, 1, 1, 1), (-1, -1, True, 1, 0, 0, 1),
           (0, -1, True, 1, 0, 1, 1), (-1, 0, True, 1, 0, 1, 1)]
if len(sys.argv) > 2 and sys.argv[2] == "sk":
    CONFIGS = [CONFIGS[0], CONFIGS[2]]
if len(sys.argv) > 2 and sys.argv[2] == "branch":
    CONFIGS = CONFIGS[:2]
PREP = [1]
if len(sys.argv) > 2 and sys.argv[2] == "prep":      # warp-MMA vs SIMT recnet_prep, everything else at its default
    CONFIGS, PREP = CONFIGS[:1], [1, 0]
for rnd in range(10):
    order = CONFIGS[rnd % len(CONFIGS):] + CONFIGS[:rnd % len(CONFIGS)]      # rotate: no config always runs first
    for (pdl, pair, fuse, lean, sk, strip, br), prep in [(c, q) for c in order for q in (PREP if rnd % 2 == 0 else PREP[::-1])]:
        lib.ffr_debug_set_prep_mma(prep)
        lib.ffr_debug_set_pdl(pdl)
        lib.ffr_debug_set_pair(pair)
        enc.fuse_se = fuse
        lib.ffr_debug_set_lean_epilogue(lean)
        lib.ffr_debug_set_streamk(sk)
        lib.ffr_debug_set_stem_strip(strip)
        rec.branch_streams = bool(br)
        run(3)
        res.setdefault("pdl=%d pair=%d fused_se=%d lean=%d streamk=%d strip_stem=%d branch_streams=%d prep_mma=%d" %
                       (pdl, pair, int(fuse), lean, sk, strip, br, prep), []).append(run(10))
rec.branch_streams = False
lib.ffr_debug_set_prep_mma(1)
enc.fuse_se = True
lib.ffr_debug_set_lean_epilogue(1)
lib.ffr_debug_set_streamk(0)
lib.ffr_debug_set_stem_strip(1)
lib.ffr_debug_set_pdl(-1)
lib.ffr_debug_set_pair(-1)
import statistics
out = {k: {"ms": [round(v, 3) for v in vs], "best_ms": round(min(vs), 3), "median_ms": round(statistics.median(vs), 3),
           "img_s_median": round(n / statistics.median(vs) * 1e3)} for k, vs in res.items()}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_bench_%d.json" % n, "w"), indent=1)
