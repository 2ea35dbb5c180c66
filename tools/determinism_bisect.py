#!/usr/bin/env python
"""Runs the RecNet training forward + backward twice on identical inputs and reports, stage by stage, the first buffer
that is not bit-identical between the two runs (TrainEngine.trace). Also poisons freed device memory with NaN before a
second engine is built, so that a read of uninitialised workspace shows up as NaN. Developer tool."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth                      # noqa: E402
from ffr_net_b200 import recnet_train               # noqa: E402
from ffr_net_b200.recnet import RecNet              # noqa: E402


def run(eng, rec, y, G, dv, dfs, dfc):
    eng.trace = []
    ws = eng.forward(y, G)
    grads = {k: torch.zeros_like(p) for k, p in rec.named_parameters()}
    eng.backward(ws, grads, dv=dv, dfs=dfs, dfc=dfc)
    torch.cuda.synchronize()
    tr = eng.trace + [("grad." + k, g.clone()) for k, g in grads.items() if k != "classifier.weight"]
    eng.trace = None
    return tr


def compare(a, b, label):
    bad = []
    for (na, ta), (nb, tb) in zip(a, b):
        assert na == nb
        same = torch.equal(ta, tb)
        nan = bool(torch.isnan(ta.float()).any() or torch.isnan(tb.float()).any())
        if not same or nan:
            d = (ta.float() - tb.float()).abs().max().item()
            bad.append({"stage": na, "max_abs_diff": d, "nan": nan, "scale": ta.float().abs().max().item()})
    print("%s: %d stages, %d differ%s" % (label, len(a), len(bad), "" if not bad else "; first: %s" % bad[0]))
    return bad


def main():
    n, G = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 2)
    out = {}
    g = torch.Generator().manual_seed(0)
    y = (torch.randn(G * n, 512, 7, 7, generator=g) * 0.5).cuda()
    dv = (torch.randn(G * n, 512, generator=g) * 1e-3).cuda()
    dfs = torch.zeros(G * n * 81, 512).cuda()
    dfc = torch.zeros(G * n * 81, 512).cuda()
    rsd = synth.synth_recnet_state_dict(0)
    rec = RecNet()
    rec.load_state_dict(rsd)
    rec = rec.cuda().train()
    eng = recnet_train.engine(rec)
    a = run(eng, rec, y, G, dv, dfs, dfc)
    rec.load_state_dict(rsd)                             # same running statistics again
    b = run(eng, rec, y, G, dv, dfs, dfc)
    out["same_engine"] = compare(a, b, "same engine, run 1 vs run 2")
    # second engine at different addresses, freed memory poisoned with NaN first
    junk = [torch.full((64 * 1024 * 1024,), float("nan"), device="cuda") for _ in range(8)]
    del junk
    rec2 = RecNet()
    rec2.load_state_dict(rsd)
    rec2 = rec2.cuda().train()
    eng2 = recnet_train.engine(rec2)
    c = run(eng2, rec2, y, G, dv, dfs, dfc)
    out["fresh_engine_poisoned"] = compare(a, c, "fresh engine on NaN-poisoned memory vs run 1")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/determinism_bisect_n%d_g%d.json" % (n, G), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
