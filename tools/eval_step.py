"""Runs a few eval steps (IR-SE50 + RecNet embedding extraction, batch 512) — the target of the ncu launch-list pass:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/eval_step.py 3
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(synth.synth_backbone_state_dict(0))
enc = enc.cuda().eval()
rec = RecNet()
rec.load_state_dict(synth.synth_recnet_state_dict(0))
rec = rec.cuda().eval()
x = synth.synth_faces(64, seed=0).repeat((n + 63) // 64, 1, 1, 1)[:n].cuda()
with torch.no_grad():
    for _ in range(steps):
        f = rec.embed_from_images(enc, x)
torch.cuda.synchronize()
print("ok", tuple(f.shape))
