"""Developer check (GPU box): run-to-run reproducibility of the RecNet training path on identical inputs, with the
caching allocator's free memory poisoned between runs (NaN or large finite garbage), so that any read of uninitialised
workspace that reaches a result shows up as NaN / a gross difference. Not part of the test-suite contract."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone as ob
from oracle import recnet as orr
from ffr_net_b200.recnet import RecNet
from ffr_net_b200 import recnet_train as rt


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def poison(value, mb=3000):
    t = torch.full((mb * 1024 * 1024 // 4,), value, dtype=torch.float32, device="cuda")
    del t


def run_layer(cin, cout, n, with_res, seed=0):
    g = torch.Generator().manual_seed(seed)
    cin_p, cout_p = rt._ceil64(cin), rt._ceil64(cout)
    x = torch.randn(n, cin, 7, 7, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda().requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(cout, generator=g)).cuda().requires_grad_(True)
    bet = (0.1 * torch.randn(cout, generator=g)).cuda().requires_grad_(True)
    slope = torch.full((cout,), 0.25).cuda().requires_grad_(True)
    up = torch.randn(n, cout, 7, 7, generator=g).cuda()
    xr = x.clone().requires_grad_(True)
    return x, xr, w, gam, bet, slope, up


def trainer_grads(fused, feats, rsd, bsd, a, b, label):
    from ffr_net_b200.trainer import Trainer, default_opts

    class Fixed:
        def __init__(self):
            self.i = 0

        def __call__(self, x):
            self.i += 1
            return feats[(self.i - 1) % 2]

    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(fused_head=fused, merge_encoder_batches=False), recnet=rec, encoder_weights=bsd)
    if not feats:
        with torch.no_grad():
            feats.extend([tuple(t.clone() for t in tr.encoder(a)), tuple(t.clone() for t in tr.encoder(b))])
    tr.encoder = Fixed()
    tr.set_input(a, b, label)
    tr.forward()
    outs = {"f_non": tr.f_non.detach().clone(), "M_space_non": tr.M_space_non.detach().clone(),
            "M_channel_non": tr.M_channel_non.detach().clone(), "space_non": tr.space_non.detach().clone(),
            "channel_non": tr.channel_non.detach().clone(), "f_ocl": tr.f_ocl.detach().clone()}
    tr.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    return outs, {k: p.grad.clone() for k, p in rec.named_parameters()}, [float(l.detach()) for l in tr.loss_items]


def backward_twice(rsd, bsd, a, b, label, fused):
    """One forward, two backward passes over the SAME saved graph: isolates the backward kernels' reproducibility."""
    from ffr_net_b200.trainer import Trainer, default_opts
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(fused_head=fused, merge_encoder_batches=False), recnet=rec, encoder_weights=bsd)
    tr.set_input(a, b, label)
    tr.forward()
    orig = torch.Tensor.backward
    torch.Tensor.backward = lambda self, *ar, **kw: orig(self, *ar, retain_graph=True, **kw)
    try:
        gs = []
        for _ in range(2):
            tr.zero_grad()
            tr.backward()
            torch.cuda.synchronize()
            gs.append({k: p.grad.clone() for k, p in rec.named_parameters()})
    finally:
        torch.Tensor.backward = orig
    d = sorted(((rel(gs[1][k], gs[0][k]), k) for k in gs[0]), reverse=True)
    print("backward twice on one saved forward (fused_head=%s): worst %s | median %.2e" %
          (fused, ", ".join("%s %.2e" % (k, v) for v, k in d[:3]), d[len(d) // 2][0]))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    a, b = ob.synth_faces(n, seed=3).cuda(), ob.synth_faces(n, seed=3, masked=True).cuda()
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(1)).cuda()
    backward_twice(rsd, bsd, a, b, label, False)
    backward_twice(rsd, bsd, a, b, label, True)
    feats = []
    base = None
    for tag, val in (("clean", None), ("clean2", None), ("poison=NaN", float("nan")), ("poison=1e4", 1.0e4),
                     ("poison=-3", -3.0)):
        if val is not None:
            poison(val)
        outs, grads, losses = trainer_grads(False, feats, rsd, bsd, a, b, label)
        nan_o = [k for k, v in outs.items() if not torch.isfinite(v).all()]
        nan_g = [k for k, v in grads.items() if not torch.isfinite(v).all()]
        if base is None:
            base = (outs, grads)
            print(tag, "losses", losses)
            continue
        do = sorted(((rel(outs[k], base[0][k]), k) for k in outs), reverse=True)
        dg = sorted(((rel(grads[k], base[1][k]), k) for k in grads), reverse=True)
        print("%-12s losses %s" % (tag, ["%.6f" % l for l in losses]))
        print("   outputs vs clean: " + ", ".join("%s %.2e" % (k, d) for d, k in do))
        print("   grads vs clean: worst %s | median %.2e | non-finite outs %s grads %d" %
              (", ".join("%s %.2e" % (k, d) for d, k in dg[:4]), dg[len(dg) // 2][0], nan_o, len(nan_g)))


if __name__ == "__main__":
    main()
