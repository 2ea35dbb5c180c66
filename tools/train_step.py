"""Runs a few EAGER training iterations (256 pairs: backbone forwards, both RecNet calls, losses, backward, clip + Adam) —
the target of ncu passes over the training kernels:
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step.py 2
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet
from ffr_net_b200.trainer import Trainer, default_opts

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda")
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(synth.synth_backbone_state_dict(0))
rec = RecNet()
rec.load_state_dict(synth.synth_recnet_state_dict(0))
tr = Trainer(default_opts(lr=1e-4), encoder=enc, recnet=rec)
rep = (pairs + 63) // 64
a = synth.synth_faces(64, seed=1).repeat(rep, 1, 1, 1)[:pairs].to(dev)
b = synth.synth_faces(64, seed=1, masked=True).repeat(rep, 1, 1, 1)[:pairs].to(dev)
label = torch.randint(0, 10575, (pairs,)).to(dev)
for _ in range(steps):
    tr.step(a, b, label)
torch.cuda.synchronize()
print("ok", [float(x.detach()) for x in tr.loss_items])
