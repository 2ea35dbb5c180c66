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


@pytest.fixture(params=["rowmajor", "pixmajor"])
def tile_mode(request, lib):
    """Runs a test with the H9 convolutions forced to row-major tiles (128 consecutive rows, sliding window) and to
    pixel-major tiles (128 images at one pixel; what batches >= ~96 use) — ffr_debug_set_pixmajor."""
    lib.ffr_debug_set_pixmajor(1 if request.param == "pixmajor" else 0)
    yield request.param
    lib.ffr_debug_set_pixmajor(-1)


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def test_wgrad_and_dgrad_kernels(lib, tile_mode):
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
    dw = torch.full((cout, cin, 3, 3), 7.0, device="cuda")                                   # overwritten, not accumulated
    ws = torch.empty(9 * rt.wgrad_workspace_elems(cout, cin), device="cuda")
    _lib.check(lib.ffr_wgrad3x3(_lib.ptr(dz_h9), 64, _lib.ptr(x_h9), 128, 0, n, cout, cin, _lib.ptr(dw), _lib.ptr(ws),
                                _lib.stream_ptr()))
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


@pytest.mark.parametrize("cin,cout,with_res", [(128, 128, True), (192, 49, False), (1536, 512, False)])
def test_convlayer_train_function(lib, cin, cout, with_res, tile_mode):
    """One train-mode ConvLayer (+ residual) through the autograd Function vs torch autograd in fp32 on the same
    bf16-rounded input: output and all five gradients."""
    import torch.nn.functional as F
    from ffr_net_b200 import recnet_train as rt
    from ffr_net_b200.recnet import ConvLayer, _h9_scatter
    g = torch.Generator().manual_seed(cin + cout)
    n = 5
    layer = ConvLayer(cin, cout, norm_type="bn", relu_type="prelu").cuda()
    with torch.no_grad():
        layer.conv2d.weight.copy_(torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5))
        layer.norm.norm.weight.copy_(torch.empty(cout).uniform_(0.5, 1.5, generator=g))
        layer.norm.norm.bias.copy_(torch.empty(cout).uniform_(-0.3, 0.3, generator=g))
        layer.relu.func.weight.copy_(torch.empty(cout).uniform_(0.1, 0.4, generator=g))
    x = (torch.randn(n, cin, 7, 7, generator=g)).bfloat16().float()
    go = torch.randn(n, cout, 7, 7, generator=g)
    # reference (fp32, CPU)
    xr = x.clone().requires_grad_(True)
    W = layer.conv2d.weight.detach().cpu().bfloat16().float().requires_grad_(True)
    gam = layer.norm.norm.weight.detach().cpu().clone().requires_grad_(True)
    bet = layer.norm.norm.bias.detach().cpu().clone().requires_grad_(True)
    slo = layer.relu.func.weight.detach().cpu().clone().requires_grad_(True)
    z = F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), W)
    y = F.batch_norm(z, None, None, gam, bet, True, 0.1, 1e-5)
    a = F.prelu(y, slo)
    if with_res:
        a = a + xr
    a.backward(go)
    # device
    xd = x.cuda().requires_grad_(True)
    tab = _h9_scatter(0, "cuda")
    cin_p = (cin + 63) // 64 * 64
    xh = rt._NchwToH9.apply(xd, cin_p)
    oh = rt._ConvLayerTrain.apply(xh, layer.conv2d.weight, layer.norm.norm.weight, layer.norm.norm.bias,
                                  layer.relu.func.weight, xh if with_res else None, layer, tab)
    out = rt._H9ToNchw.apply(oh, cout)
    out.backward(go.cuda())
    torch.cuda.synchronize()
    res = {"out": (out.detach().cpu(), a.detach()), "dx": (xd.grad.cpu(), xr.grad),
           "dW": (layer.conv2d.weight.grad.cpu(), W.grad), "dgamma": (layer.norm.norm.weight.grad.cpu(), gam.grad),
           "dbeta": (layer.norm.norm.bias.grad.cpu(), bet.grad), "dslope": (layer.relu.func.weight.grad.cpu(), slo.grad)}
    for k, (got, ref) in res.items():
        e = rel_l2(got, ref)
        print("convlayer %s rel L2 %.3e" % (k, e))
        assert e <= 2e-2, k


def test_fused_clip_adam_matches_torch(lib):
    """ffr_clip_adam == clip_grad_value_(1.0) + torch.optim.Adam over several steps (fp32, <= 1e-6 relative)."""
    from ffr_net_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(512, 1536, 3, 3), (49,), (10575, 512), (32, 561), (1,)]
    pa = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedClipAdam(pa, lr=0.1, betas=(0.9, 0.999), weight_decay=0.01, clip_value=1.0)
    ob_ = torch.optim.Adam(pb, lr=0.1, betas=(0.9, 0.999), weight_decay=0.01)
    sch = torch.optim.lr_scheduler.MultiStepLR(oa, [2], gamma=0.5)
    schb = torch.optim.lr_scheduler.MultiStepLR(ob_, [2], gamma=0.5)
    for it in range(4):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).cuda() * 3
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        torch.nn.utils.clip_grad_value_(pb, 1.0)
        ob_.step()
        sch.step()
        schb.step()
        for x, y in zip(pa, pb):
            assert rel_l2(x.detach().cpu(), y.detach().cpu()) <= 1e-6
            assert torch.equal(x.grad, y.grad)           # clipped gradients are written back


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


def test_train_step_gradients_match_oracle(lib, tile_mode):
    """Full Trainer.forward + backward (2 encoder fwd, 2 RecNet fwd with label, 4 losses, backward).

    Two comparisons, RecNet fed with the oracle's backbone outputs so that only RecNet + losses are under test:
      (a) against the pure fp32 oracle: losses within 1e-3 relative; gradient tensors within 0.35 relative L2 (median
          <= 0.15);
      (b) against the fp32 oracle with bf16 STORAGE emulated on the CPU (ConvLayer inputs, weights, raw conv outputs
          and the gradients through them rounded to bf16): within 0.3 (median <= 0.15; measured over repeated runs: worst
          0.16-0.19, median 0.093-0.096 — the bounds leave room for the run-to-run noise of the device step).
    Why so loose when every kernel is within 2e-2 in isolation (test_convlayer_train_function, 5 gradients x 3
    shapes)? At batch 4 the BatchNorm backward subtracts a large common mode (the pooled-feature gradient is constant
    over the 49 pixels of a sample), which amplifies the 2^-9 rounding of bf16-stored gradients layer after layer:
    the CPU emulation alone deviates from pure fp32 by the SAME profile (worst 0.25 / median 0.10, growing with
    backward depth: classifier 5e-3 -> Conv4Merge 6e-2 -> ChannelFlipMerge 1.2e-1 -> Conv4Space 2e-1). The device
    path and the emulation are two noisy realisations of an ill-conditioned map, not bit-identical roundings."""
    from ffr_net_b200.backbone import Backbone
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=5), ob.synth_faces(n, seed=5, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(5))
    items_ref, grads_ref, stats_ref, _ = otr.train_step(bsd, rsd, img1, img2, label)
    items_emu, grads_emu, _, _ = otr.train_step(bsd, rsd, img1, img2, label, emulate_bf16=True)
    enc, rec = Backbone(50, 0.6, "ir_se"), RecNet()
    enc.load_state_dict(bsd)
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(), encoder=enc, recnet=rec)
    with torch.no_grad():
        y1, e1 = ob.backbone_forward(bsd, img1)
        y2, e2 = ob.backbone_forward(bsd, img2)
    tr.gt_label = label.cuda()
    tr.feat_map_non, tr.feat_extract_non = y1.cuda(), e1.cuda()
    tr.feat_map_ocl, tr.feat_extract_ocl = y2.cuda(), e2.cuda()
    (tr.f_non, tr.pred_loss_non, tr.pred_label_non, tr.M_space_non, tr.M_channel_non, tr.space_non,
     tr.channel_non) = rec(tr.feat_map_non, tr.gt_label)
    (tr.f_ocl, tr.pred_loss_ocl, tr.pred_label_ocl, tr.M_space_ocl, tr.M_channel_ocl, tr.space_ocl,
     tr.channel_ocl) = rec(tr.feat_map_ocl, tr.gt_label)
    tr.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    items = [float(v) for v in tr.loss_items]
    print("losses", items, "fp32 oracle", items_ref)
    for a, b in zip(items, items_ref):
        assert abs(a - b) <= 1e-3 * max(abs(b), 1e-3)
    named = dict(rec.named_parameters())
    assert set(named) == set(grads_ref) and len(named) == 76
    e_emu = sorted(((rel_l2(p.grad.cpu(), grads_emu[k]), k) for k, p in named.items()), reverse=True)
    e_f32 = sorted(((rel_l2(p.grad.cpu(), grads_ref[k]), k) for k, p in named.items()), reverse=True)
    print("vs bf16-storage emulation: worst %.3e (%s) median %.3e" % (e_emu[0][0], e_emu[0][1], e_emu[38][0]))
    print("vs pure fp32 oracle:       worst %.3e (%s) median %.3e" % (e_f32[0][0], e_f32[0][1], e_f32[38][0]))
    assert e_f32[0][0] <= 0.35 and e_f32[38][0] <= 0.15, e_f32[:3]
    assert e_emu[0][0] <= 0.3 and e_emu[38][0] <= 0.15, e_emu[:3]
    cos = min(torch.nn.functional.cosine_similarity(p.grad.cpu().reshape(1, -1), grads_ref[k].reshape(1, -1)).item()
              for k, p in named.items())
    print("min gradient cosine vs fp32 oracle %.4f" % cos)
    assert cos >= 0.95
    k = "Conv4Merge.0.norm.norm.running_mean"
    assert rel_l2(rec.state_dict()[k].cpu(), stats_ref[k]) <= 2e-2
    assert int(rec.state_dict()["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 2   # two recnet calls per step
    # the whole step through the public Trainer API (bf16 backbone in the loop), then clip + Adam + LR step
    rec.load_state_dict(rsd)
    tr.set_input(img1.cuda(), img2.cuda(), label.cuda())
    tr.forward()
    tr.optimizer_parameters(0)
    tr.update_learning_rate()
    vals = tr.get_current_values()
    print("trainer step:", vals)
    assert all(torch.isfinite(p).all() for p in rec.parameters())
    assert abs(float(vals["ClassifierLoss"]) - items_ref[3]) <= 2e-2 * items_ref[3]


@pytest.mark.parametrize("split", [False, True])
def test_cuda_graph_step_matches_eager(lib, split):
    """split=True: the data-parallel form (graph 1 = forward + backward into the flat gradient buffer, [all-reduce],
    graph 2 = clip+Adam) exercised on one GPU.
    Trainer.capture_step(): ONE replay of the captured iteration from a given state equals ONE eager iteration from
    the same state (parameters, BN buffers, Adam moments, step count) up to the round-off of the fp32 atomics.
    (Multi-step trajectories are not compared: Adam's sign-like update amplifies that round-off chaotically.)"""
    from ffr_net_b200.backbone import Backbone
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=7).cuda(), ob.synth_faces(n, seed=7, masked=True).cuda()
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(7)).cuda()

    def make():
        enc, rec = Backbone(50, 0.6, "ir_se"), RecNet()
        enc.load_state_dict(bsd)
        rec.load_state_dict(rsd)
        return Trainer(default_opts(lr=1e-3), encoder=enc, recnet=rec)
    eager, graphed = make(), make()
    for _ in range(2):
        eager.step(img1, img2, label)
    graphed.capture_step(img1, img2, label, warmup=2, split_optimizer=split)
    assert (graphed._graph_opt is not None) == split and graphed._flat_bound
    # put `graphed` into exactly the state of `eager` (in place: the graph holds the tensor addresses)
    with torch.no_grad():
        for a, b in zip(graphed.recnet.parameters(), eager.recnet.parameters()):
            a.copy_(b)
            sa, sb = graphed.optim.state[a], eager.optim.state[b]
            sa["exp_avg"].copy_(sb["exp_avg"])
            sa["exp_avg_sq"].copy_(sb["exp_avg_sq"])
        for a, b in zip(graphed.recnet.buffers(), eager.recnet.buffers()):
            a.copy_(b)
        graphed.optim._tables[0][4].copy_(eager.optim._tables[0][4])
    before = [p.detach().clone() for p in eager.recnet.parameters()]
    eager.step(img1, img2, label)
    graphed.step(img1, img2, label)
    torch.cuda.synchronize()
    # compare the UPDATES (parameter deltas), relative to the update size
    num = den = 0.0
    for a, b, p0 in zip(graphed.recnet.parameters(), eager.recnet.parameters(), before):
        num += float(((a.detach() - b.detach()).double() ** 2).sum())
        den += float(((b.detach() - p0).double() ** 2).sum())
    rel = (num / den) ** 0.5
    print("graph replay vs eager step: relative difference of the parameter update %.3e" % rel)
    assert den > 0 and rel <= 1e-1          # measured 2.9e-2 .. 3.3e-2 over repeated runs (bf16 rounding-noise realisations)
    assert float(graphed.optim._tables[0][4][1]) == float(eager.optim._tables[0][4][1]) == 3.0
    k = "Conv4Merge.0.norm.norm.num_batches_tracked"
    assert int(graphed.recnet.state_dict()[k]) == int(eager.recnet.state_dict()[k]) == 6


def test_trainer_checkpoint_roundtrip(lib, tmp_path):
    """Trainer.save_model / load_model (models/trainer.py:201-224): `<ckpt_dir>/<name>.pth.gzip` with the reference's
    container keys; a fresh trainer that loads it continues from the same weights (the packed bf16 caches are rebuilt)."""
    from ffr_net_b200 import checkpoint as ck
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    a, b = ob.synth_faces(4, seed=3).cuda(), ob.synth_faces(4, seed=3, masked=True).cuda()
    label = torch.tensor([5, 17, 10000, 3], device="cuda")
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(lr=1e-3, ckpt_dir=str(tmp_path)), recnet=rec, encoder_weights=bsd)
    tr.step(a, b, label)
    path = tr.save_model("epoch_000_iter_000001", {"epoch": 0, "iter": 1})
    w = ck.load(path, map_location="cpu")
    assert set(w.keys()) == {"RecNet", "optimizer", "epoch", "iter"} and len(w["RecNet"]) == 121
    rec2 = RecNet()
    tr2 = Trainer(default_opts(lr=1e-3, ckpt_dir=str(tmp_path)), recnet=rec2, encoder_weights=bsd)
    tr2.load_model("latest")
    assert tr2.start_point == {"epoch": 0, "iter": 1}
    for (k, p), (_, q) in zip(rec.state_dict().items(), rec2.state_dict().items()):
        assert torch.equal(p, q), k
    tr.recnet.eval()
    tr2.recnet.eval()
    with torch.no_grad():
        y, _ = tr.encoder(a)
        v1, _ = tr.recnet(y)
        v2, _ = tr2.recnet(y)
    assert (v1 - v2).abs().max().item() <= 1e-5 * v1.abs().max().item() + 1e-6


def test_two_stream_recnet_calls_match_sequential(lib):
    """opts.two_streams: the masked RecNet call on a side stream, concurrent with the unmasked one. Same losses,
    BatchNorm running statistics applied in the reference's order (two updates per layer), gradients equal up to the
    run-to-run noise of the sequential step; and the whole thing replays from a CUDA graph."""
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 8
    a, b = ob.synth_faces(n, seed=5).cuda(), ob.synth_faces(n, seed=5, masked=True).cuda()
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(5)).cuda()
    res = []
    for two in (False, False, True):
        rec = RecNet()
        rec.load_state_dict(rsd)
        tr = Trainer(default_opts(lr=1e-3, two_streams=two), recnet=rec, encoder_weights=bsd)
        tr.set_input(a, b, label)
        tr.forward()
        tr.zero_grad()
        tr.backward()
        torch.cuda.synchronize()
        res.append(([float(l.detach()) for l in tr.loss_items], {k: p.grad.clone() for k, p in rec.named_parameters()},
                    {k: v.clone() for k, v in rec.state_dict().items() if "running" in k or "tracked" in k}))
    (l0, g0, s0), (l0b, g0b, _), (l1, g1, s1) = res
    assert all(abs(x - y) <= 1e-3 * max(1.0, abs(x)) for x, y in zip(l0, l1)), (l0, l1)
    for k in s0:
        if k.endswith("num_batches_tracked"):
            assert int(s0[k]) == int(s1[k]) == 2
        else:
            assert rel_l2(s1[k], s0[k]) <= 2e-3, k
    cosd = sorted(torch.nn.functional.cosine_similarity(g1[k].reshape(1, -1), g0[k].reshape(1, -1)).item() for k in g0)
    cosn = sorted(torch.nn.functional.cosine_similarity(g0b[k].reshape(1, -1), g0[k].reshape(1, -1)).item() for k in g0)
    print("two-stream vs sequential grad cosine: min %.4f median %.4f | run-to-run: min %.4f median %.4f" %
          (cosd[0], cosd[len(cosd) // 2], cosn[0], cosn[len(cosn) // 2]))
    assert cosd[len(cosd) // 2] >= min(0.99, cosn[len(cosn) // 2] - 0.01) and cosd[0] >= cosn[0] - 0.1
    # graph capture + a few replays with the side stream inside the graph
    rec = RecNet()
    rec.load_state_dict(rsd)
    for split in (False, True):              # one graph, and the data-parallel form (two graphs around the exchange)
        rec = RecNet()
        rec.load_state_dict(rsd)
        tr = Trainer(default_opts(lr=1e-3, two_streams=True), recnet=rec, encoder_weights=bsd)
        tr.capture_step(a, b, label, warmup=2, split_optimizer=split)
        for _ in range(3):
            tr.step(a, b, label)
        torch.cuda.synchronize()
        vals = tr.get_current_values()
        assert all(torch.isfinite(p).all() for p in rec.parameters())
        assert abs(float(vals["SelfSimilarityLoss"]) - l0[0]) <= 0.1 * max(1.0, abs(l0[0]))
        # 2 calls per iteration x (2 warm-up iterations + 3 replays); the capture itself executes nothing
        assert int(rec.state_dict()["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 2 * (2 + 3)
