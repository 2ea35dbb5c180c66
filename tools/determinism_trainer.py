#!/usr/bin/env python
"""Runs Trainer.forward + backward on identical (precomputed) backbone features with two fresh Trainers and reports every
traced buffer that is not bit-identical between them: engine stages (TrainEngine.trace), loss / head buffers, gradients.
Developer tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth, recnet_train          # noqa: E402
from ffr_net_b200.recnet import RecNet               # noqa: E402
from ffr_net_b200.trainer import Trainer, default_opts   # noqa: E402


def run(n, feats, label, rsd, bsd):
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(), recnet=rec, encoder_weights=bsd)
    tr.encoder = lambda x: feats
    eng = recnet_train.engine(rec)
    eng.trace = []
    tr.set_input(torch.zeros(n, 3, 112, 112, device="cuda"), torch.zeros(n, 3, 112, 112, device="cuda"), label)
    tr.forward()
    lw = tr._lw
    t = [("fwd." + k, getattr(lw.head, k).clone()) for k in ("vp", "cos", "sumexp", "zlabel", "argkey")]
    t.append(("fwd.v", lw.v.clone()))
    t.append(("fwd.ce", lw.ce.clone()))
    mark = len(eng.trace)
    tr.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    t += [("loss." + k, getattr(lw, k).clone()) for k in ("A6", "B6", "D", "spart", "chan_sums", "e", "dfc", "dfs",
                                                            "space_part", "row_part", "dvl", "dv", "out")]
    t += [("head." + k, getattr(lw.head, k).clone()) for k in ("dcos", "dcos_t", "dvh", "dv", "dwh")]
    own = torch.zeros(81, dtype=torch.bool)
    own[[(h + 1) * 9 + (w + 1) for h in range(7) for w in range(7)]] = True
    eng_tr = []
    for name, ten in eng.trace:
        if name.endswith(".afold") or name == "fwd.prep.cm":     # only the own rows are defined
            if name.endswith(".afold"):
                ten = ten.view(-1, 81, ten.shape[1])[:, own.to(ten.device)]
            else:
                continue
        eng_tr.append((name, ten))
    t += eng_tr
    t += [("grad." + k, p.grad.clone()) for k, p in rec.named_parameters()]
    return t


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    g = torch.Generator().manual_seed(0)
    y = (torch.randn(2 * n, 512, 7, 7, generator=g) * 0.5).cuda()
    f = torch.nn.functional.normalize(torch.randn(2 * n, 512, generator=g)).cuda()
    label = torch.randint(0, 10575, (n,), generator=g).cuda()
    rsd, bsd = synth.synth_recnet_state_dict(0), synth.synth_backbone_state_dict(0)
    a = run(n, (y, f), label, rsd, bsd)
    b = run(n, (y, f), label, rsd, bsd)
    bad = []
    for (na, ta), (nb, tb) in zip(a, b):
        assert na == nb, (na, nb)
        if not torch.equal(ta, tb):
            d = (ta.double() - tb.double())
            bad.append({"stage": na, "max_abs_diff": d.abs().max().item(), "rel_l2": (d.norm() / (ta.double().norm() + 1e-30)).item(),
                        "n_diff": int((ta != tb).sum())})
    print("n=%d: %d stages, %d differ" % (n, len(a), len(bad)))
    for x in bad[:25]:
        print("  ", x)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/determinism_trainer_n%d.json" % n, "w") as fjs:
        json.dump(bad, fjs, indent=1)


if __name__ == "__main__":
    main()
