#!/usr/bin/env python
"""Exploration (developer tool): how many fitting steps a randomly initialised RecNet needs before the rectified embeddings
of the synthetic verification set spread (ffr_net_b200/lfw.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import lfw, synth                          # noqa: E402
from ffr_net_b200.recnet import RecNet                      # noqa: E402
from ffr_net_b200.trainer import Trainer, default_opts      # noqa: E402

out = []
bsd, rsd = synth.synth_backbone_state_dict(0), synth.synth_recnet_state_dict(0)
img = lfw.synth_pairs(0, 600, 0, 600)
for lr, steps in ((1e-3, 0), (1e-3, 50), (1e-3, 200), (3e-3, 200)):
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(lr=lr), recnet=rec, encoder_weights=bsd)
    lfw.fit_recnet(tr, steps=steps)
    vals = tr.get_current_values() if steps else {}
    rec.eval()
    r = lfw.verify(tr.encoder, rec, n_pairs=600, images=img)
    sn, sr, lab = r["scores_rectified"], r["scores_raw"], r["labels"].bool()
    row = {"lr": lr, "steps": steps, "acc_rect": r["acc_rectified"], "acc_raw": r["acc_raw"],
           "rect_same": [sn[lab].min().item(), sn[lab].mean().item()], "rect_diff": [sn[~lab].mean().item(), sn[~lab].max().item()],
           "raw_same": [sr[lab].min().item(), sr[lab].mean().item()], "raw_diff": [sr[~lab].mean().item(), sr[~lab].max().item()],
           "losses": vals}
    print(json.dumps(row))
    out.append(row)
    rec.train()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lfw_explore.json", "w"), indent=1)
