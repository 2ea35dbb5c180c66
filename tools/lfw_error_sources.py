#!/usr/bin/env python
"""Where does the cosine error of the rectified embeddings come from? (developer tool) Compares, on 200 synthetic pairs and
a briefly fitted RecNet: fp32 oracle end to end | fp32 oracle RecNet on the DEVICE backbone's feature maps | device RecNet
(bf16 eval path) on the ORACLE's feature maps | device end to end."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone as ob, recnet as orr, scoring as osc      # noqa: E402  (developer tool, not product)
from ffr_net_b200 import lfw                                         # noqa: E402
from ffr_net_b200.recnet import RecNet                              # noqa: E402
from ffr_net_b200.trainer import Trainer, default_opts              # noqa: E402

N = 200
bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
rec = RecNet()
rec.load_state_dict(rsd)
tr = Trainer(default_opts(lr=1e-3), recnet=rec, encoder_weights=bsd)
lfw.fit_recnet(tr, steps=50, batch=32)
rec.eval()
fitted = {k: v.detach().cpu().clone() for k, v in rec.state_dict().items()}
img1, img2 = lfw.synth_pairs(0, N, 0, 600)
with torch.no_grad():
    yr = [ob.backbone_forward(bsd, im) for im in (img1, img2)]
    yd = [tuple(t.cpu() for t in tr.encoder(im.cuda())) for im in (img1, img2)]
    v_ref = [orr.recnet_forward(fitted, y)[0] for y, _ in yr]
    v_ydev = [orr.recnet_forward(fitted, y)[0] for y, _ in yd]
    v_dev_yref = [rec(y.cuda())[0].cpu() for y, _ in yr]
    v_dev = [rec(y.cuda())[0].cpu() for y, _ in yd]
cos = lambda a, b: osc.pair_cosine(a, b).numpy()
c_ref = cos(*v_ref)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
out = {
    "feature_map_rel_err_device_backbone": rel(yd[0][0], yr[0][0]),
    "raw_embedding_rel_err": rel(yd[0][1], yr[0][1]),
    "raw_cos_err": float(np.abs(cos(yd[0][1], yd[1][1]) - cos(yr[0][1], yr[1][1])).max()),
    "rect_embedding_rel_err": {"oracle_recnet_on_device_y": rel(v_ydev[0], v_ref[0]), "device_recnet_on_oracle_y": rel(v_dev_yref[0], v_ref[0]),
                               "device_end_to_end": rel(v_dev[0], v_ref[0])},
    "rect_cos_err_max": {"oracle_recnet_on_device_y": float(np.abs(cos(*v_ydev) - c_ref).max()),
                         "device_recnet_on_oracle_y": float(np.abs(cos(*v_dev_yref) - c_ref).max()),
                         "device_end_to_end": float(np.abs(cos(*v_dev) - c_ref).max())},
    "rect_cos_err_median": {"oracle_recnet_on_device_y": float(np.median(np.abs(cos(*v_ydev) - c_ref))),
                            "device_recnet_on_oracle_y": float(np.median(np.abs(cos(*v_dev_yref) - c_ref))),
                            "device_end_to_end": float(np.median(np.abs(cos(*v_dev) - c_ref)))},
}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/lfw_error_sources.json", "w"), indent=1)
