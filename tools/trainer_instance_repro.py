"""Are the gradients of a training step reproducible across Trainer INSTANCES (same weights, same data, same GPU)?
Compares the frozen-backbone features and every gradient of two trainers, and of one trainer run twice."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth                               # noqa: E402
from ffr_net_b200.recnet import RecNet                       # noqa: E402
from ffr_net_b200.trainer import Trainer, default_opts       # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    bsd, rsd = synth.synth_backbone_state_dict(0), synth.synth_recnet_state_dict(0)
    a = synth.synth_faces(n, seed=100).cuda()
    b = synth.synth_faces(n, seed=100, masked=True).cuda()
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(100)).cuda()

    def make():
        rec = RecNet()
        rec.load_state_dict(rsd)
        return Trainer(default_opts(lr=1e-3, data_parallel=False), recnet=rec, encoder_weights=bsd)

    def run(tr):
        tr.set_input(a, b, label)
        tr.forward()
        tr.zero_grad()
        tr.backward()
        torch.cuda.synchronize()
        g = {k: p.grad.clone() for k, p in tr.recnet.named_parameters()}
        feats = {"y": tr._y.clone(), "feat_extract_non": tr.feat_extract_non.clone(), "v": tr._lw.v.clone()}
        return g, feats

    t1, t2 = make(), make()
    g1, f1 = run(t1)
    g1b, f1b = run(t1)
    g2, f2 = run(t2)
    out = {"n": n}
    for name, (ga, fa, gb, fb) in {"same_instance": (g1, f1, g1b, f1b), "two_instances": (g1, f1, g2, f2)}.items():
        e = {k: ((ga[k].double() - gb[k].double()).norm() / (ga[k].double().norm() + 1e-30)).item() for k in ga}
        out[name] = {"grads_equal": all(torch.equal(ga[k], gb[k]) for k in ga),
                     "worst": sorted(e.items(), key=lambda kv: -kv[1])[:3],
                     "features": {k: {"equal": bool(torch.equal(fa[k], fb[k])),
                                      "max_abs": (fa[k].float() - fb[k].float()).abs().max().item(),
                                      "n_diff": int((fa[k] != fb[k]).sum())} for k in fa}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
