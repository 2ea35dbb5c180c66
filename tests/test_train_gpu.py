"""GPU parity of the training path (RecNet train-mode forward, backward kernels, full Trainer step) against the fp32
CPU oracle. Activations and activation-gradients are bf16 on the device, so tolerances are relative L2 errors:
outputs <= 2e-2, parameter gradients <= 8e-2 per tensor (stated per assert)."""
import numpy as np
import pytest
import torch

from oracle import backbone as ob
from oracle import recnet as orr
from oracle import train as otr
from ffr_net_b200 import _lib
from ffr_net_b200.recnet import RecNet

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def test_wgrad_and_dgrad_kernels(lib):
    """ffr_wgrad3x3 and the flipped-weight dgrad against autograd of F.conv2d(reflect-padded) on bf16 operands."""
    import torch.nn.functional as F
    from ffr_net_b200 import recnet_train as rt
    g = torch.Generator().manual_seed(0)
    n, cin, cout = 3, 128, 64
    x = torch.randn(n, cin, 7, 7, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    dz = torch.randn(n, cout, 7, 7, generator=g)
    xb, wb, dzb = x.bfloat16().float(), w.bfloat16().float(), dz.bfloat16().float()
    xr, wr = xb.clone().requires_grad_(True), wb.clone().requires_grad_(True)
    z = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), wr)
    z.backward(dzb)
    x_h9 = rt._NchwToH9.apply(x.cuda(), 128)
    dz_h9 = rt._H9ToNchw.backward(type("c", (), {"dims": (n, cout, 64)}), dz.cuda())[0]       # zero-halo H9
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    _lib.check(lib.ffr_wgrad3x3(_lib.ptr(dz_h9), 64, _lib.ptr(x_h9), 128, 0, n, cout, cin, _lib.ptr(dw), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert rel_l2(dw.cpu(), wr.grad) <= 5e-3
    wt = torch.zeros(128, 3, 3, 64, dtype=torch.bfloat16, device="cuda")
    wt[:cin, :, :, :cout] = w.cuda().flip(2, 3).permute(1, 2, 3, 0).to(torch.bfloat16)
    dx = torch.empty(n * 81, 128, dtype=torch.bfloat16, device="cuda")
    rt._conv_gemm(lib, dz_h9, wt.reshape(128, 9 * 64), 64, 128, n * 81, n, 0, dx)
    dx_nchw = torch.empty(n, cin, 7, 7, device="cuda")
    _lib.check(lib.ffr_h9_to_nchw(_lib.ptr(dx), 128, 0, _lib.ptr(dx_nchw), n, cin, 1, _lib.stream_ptr()))   # fold mirrors
    torch.cuda.synchronize()
    assert rel_l2(dx_nchw.cpu(), xr.grad) <= 1e-2


@pytest.fixture(scope="module")
def models(lib):
    rsd = orr.synth_recnet_state_dict(0)
    m = RecNet()
    m.load_state_dict(rsd)
    return rsd, m.cuda().train()


def test_recnet_train_forward_matches_oracle(models):
    rsd, m = models
    m.load_state_dict(rsd)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 512, 7, 7, generator=g) * 0.3
    label = torch.randint(0, 10575, (4,), generator=g)
    with torch.no_grad():
        ref, stats = orr.recnet_forward(rsd, x, label, training=True, return_stats=True)
        out = m(x.cuda(), label.cuda())
    names = ["feat_new_v", "pred_loss", "pred_label", "M_space", "M_channel", "feat_space", "feat_channel"]
    for nme, a, b in zip(names, out, ref):
        e = rel_l2(a.cpu(), b)
        print("train fwd %-12s rel L2 %.3e" % (nme, e))
        assert a.shape == b.shape and e <= 2e-2, nme
    sd = m.state_dict()
    for k in ("Conv4Merge.0.norm.norm.running_mean", "Conv4Space.0.norm.norm.running_var"):
        assert rel_l2(sd[k].cpu(), stats[k]) <= 2e-2, k
    assert int(sd["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 1


def test_train_step_gradients_match_oracle(lib):
    """Full Trainer.forward + backward (2 encoder fwd, 2 RecNet fwd with label, 4 losses, backward): losses and all
    76 gradient tensors vs the fp32 CPU oracle."""
    from ffr_net_b200.backbone import Backbone
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=5), ob.synth_faces(n, seed=5, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(5))
    items_ref, grads_ref, stats_ref, acc_ref = otr.train_step(bsd, rsd, img1, img2, label)
    enc, rec = Backbone(50, 0.6, "ir_se"), RecNet()
    enc.load_state_dict(bsd)
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(), encoder=enc, recnet=rec)
    tr.set_input(img1.cuda(), img2.cuda(), label.cuda())
    tr.forward()
    tr.optim.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    items = [float(v) for v in tr.loss_items]
    print("losses", items, "ref", items_ref)
    for a, b in zip(items, items_ref):
        assert abs(a - b) <= 2e-2 * max(abs(b), 1e-3)
    worst = ("", 0.0)
    named = dict(rec.named_parameters())
    assert set(named) == set(grads_ref) and len(named) == 76
    for k, p in named.items():
        assert p.grad is not None, k
        e = rel_l2(p.grad.cpu(), grads_ref[k])
        if e > worst[1]:
            worst = (k, e)
    print("worst gradient rel L2: %s %.3e" % worst)
    assert worst[1] <= 8e-2, worst
    k = "Conv4Merge.0.norm.norm.running_mean"
    assert rel_l2(rec.state_dict()[k].cpu(), stats_ref[k]) <= 2e-2
    assert int(rec.state_dict()["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 2   # two recnet calls per step
    tr.allreduce_gradients()
    torch.nn.utils.clip_grad_value_(rec.parameters(), 1.0)
    tr.optim.step()
    tr.update_learning_rate()
    assert all(torch.isfinite(p).all() for p in rec.parameters())
