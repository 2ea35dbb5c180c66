"""GPU parity of the training path end to end (RecNet train-mode forward, backward, losses, optimizer, CUDA-graph replay)
against the fp32 CPU oracle (oracle/train.py, pinned to the real models/trainer.py by tests/golden). Per-kernel parity
lives in test_train_kernels_gpu.py. Tolerances (stated per assert): forward outputs <= 5e-3 relative L2 (fp16 hi+lo
activations, fp16 weights; M_channel is exported from its bf16 copy), losses <= 1e-3 relative, parameter gradients <= 5e-2
per tensor and <= 2e-2 median relative L2 against the PURE fp32 oracle at 4 and at 32 pairs, and bit-identical results
from run to run and between eager execution and CUDA-graph replay (every reduction has a fixed order)."""
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


def test_fused_clip_adam_state_dict_roundtrip(lib):
    """state_dict() has torch.optim.Adam's layout (step / exp_avg / exp_avg_sq) and load_state_dict() restores the
    moments, the step count and the learning rate into the buffers the kernel uses: a resumed optimizer continues
    exactly like the original, and the state also loads into torch.optim.Adam."""
    from ffr_net_b200.optim import FusedClipAdam
    g = torch.Generator().manual_seed(1)
    shapes = [(64, 33), (7,), (300, 5, 3, 3)]
    mk = lambda: [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(9)).cuda()) for s in shapes]
    pa, pb, pc = mk(), mk(), mk()
    oa = FusedClipAdam(pa, lr=0.05, clip_value=1.0)
    grads = [[torch.randn(s, generator=g).cuda() for s in shapes] for _ in range(5)]
    for it in range(3):
        for p, gr in zip(pa, grads[it]):
            p.grad = gr.clone()
        oa.step()
    import copy
    sd = copy.deepcopy(oa.state_dict())         # as a checkpoint round trip would (load_state_dict shares tensors)
    assert "_lr_on_device" not in sd["param_groups"][0]
    assert all(float(st["step"]) == 3.0 for st in sd["state"].values())
    with torch.no_grad():
        for q, p in zip(pb, pa):
            q.copy_(p)
        for q, p in zip(pc, pa):
            q.copy_(p)
    ob_ = FusedClipAdam(pb, lr=0.05, clip_value=1.0)
    ob_.load_state_dict(copy.deepcopy(sd))
    oc = torch.optim.Adam(pc, lr=0.05)
    oc.load_state_dict(copy.deepcopy(sd))
    for it in range(3, 5):
        for x, y, z, gr in zip(pa, pb, pc, grads[it]):
            x.grad, y.grad, z.grad = gr.clone(), gr.clone(), gr.clone()
        oa.step()
        ob_.step()
        torch.nn.utils.clip_grad_value_(pc, 1.0)
        oc.step()
    for x, y, z in zip(pa, pb, pc):
        assert torch.equal(x.detach(), y.detach())
        assert rel_l2(z.detach().cpu(), x.detach().cpu()) <= 1e-6
    assert ob_.device_step() == 5


@pytest.fixture(scope="module")
def models(lib):
    rsd = orr.synth_recnet_state_dict(0)
    m = RecNet()
    m.load_state_dict(rsd)
    return rsd, m.cuda().train()


def test_recnet_train_forward_matches_oracle(models):
    """Public training-mode RecNet.forward(x, label): the reference's 7-tuple, BatchNorm running statistics."""
    rsd, m = models
    m.load_state_dict(rsd)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 512, 7, 7, generator=g) * 0.3
    label = torch.randint(0, 10575, (4,), generator=g)
    with torch.no_grad():
        ref, stats = orr.recnet_forward(rsd, x, label, training=True, return_stats=True)
    out = m(x.cuda(), label.cuda())
    names = ["feat_new_v", "pred_loss", "pred_label", "M_space", "M_channel", "feat_space", "feat_channel"]
    tol = {"M_channel": 5e-3}
    for nme, a, b in zip(names, out, ref):
        e = rel_l2(a.detach().cpu(), b)
        print("train fwd %-12s rel L2 %.3e" % (nme, e))
        assert a.shape == b.shape and e <= tol.get(nme, 2e-3), nme
    sd = m.state_dict()
    for k in ("Conv4Merge.0.norm.norm.running_mean", "Conv4Space.0.norm.norm.running_var",
              "Conv4Space.5.conv2.norm.norm.running_var"):
        assert rel_l2(sd[k].cpu(), stats[k]) <= 2e-3, k
    assert int(sd["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 1
    v2, fmap = m(x.cuda())                                           # label=None -> 2-tuple (recnet.py:426)
    assert tuple(fmap.shape) == (4, 512, 7, 7) and rel_l2(fmap.mean((2, 3)).cpu(), v2.cpu()) <= 1e-5


def _trainer_on_oracle_features(bsd, rsd, img1, img2, label, **opts):
    """Trainer whose RecNet is fed with the ORACLE's backbone outputs, so only RecNet + losses are under test."""
    from ffr_net_b200.backbone import Backbone
    from ffr_net_b200.trainer import Trainer, default_opts
    enc, rec = Backbone(50, 0.6, "ir_se"), RecNet()
    enc.load_state_dict(bsd)
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(**opts), encoder=enc, recnet=rec)
    with torch.no_grad():
        y1, e1 = ob.backbone_forward(bsd, img1)
        y2, e2 = ob.backbone_forward(bsd, img2)
    feats = (torch.cat((y1, y2)).cuda(), torch.cat((e1, e2)).cuda())
    tr.encoder = lambda x: feats
    tr.set_input(img1.cuda(), img2.cuda(), label.cuda())
    tr.forward()
    tr.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    return tr, rec


@pytest.mark.parametrize("n", [4, 32])
def test_train_step_gradients_match_oracle(lib, n):
    """Trainer.forward + backward (2 RecNet calls with label, 4 losses, all 76 gradients) against the PURE fp32 oracle.
    n = 4: two sequential calls (G = 1 each); n = 32: the two calls batched (G = 2, per-call BatchNorm statistics)."""
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    img1, img2 = ob.synth_faces(n, seed=5), ob.synth_faces(n, seed=5, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(5))
    items_ref, grads_ref, stats_ref, acc_ref = otr.train_step(bsd, rsd, img1, img2, label)
    tr, rec = _trainer_on_oracle_features(bsd, rsd, img1, img2, label)
    assert len(tr._calls) == (2 if n % 32 else 1)
    items = [float(v) for v in tr.loss_items]
    print("losses", items, "fp32 oracle", items_ref)
    for a, b in zip(items, items_ref):
        assert abs(a - b) <= 1e-3 * max(abs(b), 1e-3)
    named = dict(rec.named_parameters())
    assert set(named) == set(grads_ref) and len(named) == 76
    e = sorted(((rel_l2(p.grad.cpu(), grads_ref[k]), k) for k, p in named.items()), reverse=True)
    print("n=%d vs pure fp32 oracle: worst %.3e (%s) median %.3e" % (n, e[0][0], e[0][1], e[38][0]))
    for err, k in e[:8]:
        print("   %.3e %s" % (err, k))
    assert e[0][0] <= 5e-2 and e[38][0] <= 2e-2, e[:3]
    for k in ("Conv4Merge.0.norm.norm.running_mean", "Conv4Space.0.norm.norm.running_var"):
        assert rel_l2(rec.state_dict()[k].cpu(), stats_ref[k]) <= 2e-3, k
    assert int(rec.state_dict()["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 2   # two recnet calls per step
    assert abs(float(tr._correct) / n - acc_ref) <= 1.0 / n
    # run-to-run: bit-identical losses and gradients (fixed-order reductions everywhere)
    tr2, rec2 = _trainer_on_oracle_features(bsd, rsd, img1, img2, label)
    assert [float(v) for v in tr2.loss_items] == items
    for (k, p), (_, q) in zip(rec.named_parameters(), rec2.named_parameters()):
        assert torch.equal(p.grad, q.grad), k


def test_literal_trainer_matches_fused(lib):
    """opts.literal: the reference's own sequence (two public RecNet calls returning 7-tuples, ATen losses under
    autograd, loss.backward()) vs the fused engine path: same losses, same gradients (the literal losses are fp32; the
    fused channel-similarity gradient passes through bf16 GEMM operands)."""
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=8), ob.synth_faces(n, seed=8, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(8))
    tf, rf = _trainer_on_oracle_features(bsd, rsd, img1, img2, label)
    tl, rl = _trainer_on_oracle_features(bsd, rsd, img1, img2, label, literal=True)
    lf, ll = [float(v) for v in tf.loss_items], [float(v.detach()) for v in tl.loss_items]
    print("fused", lf, "literal", ll)
    assert all(abs(a - b) <= 5e-4 * max(1.0, abs(b)) for a, b in zip(lf, ll))
    e = sorted(((rel_l2(p.grad, dict(rl.named_parameters())[k].grad), k) for k, p in rf.named_parameters()), reverse=True)
    print("fused vs literal gradients: worst %.3e (%s) median %.3e" % (e[0][0], e[0][1], e[38][0]))
    assert e[0][0] <= 3e-2 and e[38][0] <= 1e-2


def test_cuda_graph_step_bit_equals_eager_and_eval_sees_new_weights(lib):
    """capture_step(): (1) capturing leaves parameters, Adam state and BatchNorm buffers untouched; (2) replays equal eager
    steps bit for bit; (3) an eval-mode forward between replays uses the CURRENT weights (packed-weight caches are
    invalidated after every replay)."""
    from ffr_net_b200.backbone import Backbone
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=7).cuda(), ob.synth_faces(n, seed=7, masked=True).cuda()
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(7)).cuda()

    def make():                                   # the frozen backbone is in the loop: it is bit-reproducible too
        enc, rec = Backbone(50, 0.6, "ir_se"), RecNet()
        enc.load_state_dict(bsd)
        rec.load_state_dict(rsd)
        return Trainer(default_opts(lr=1e-3), encoder=enc, recnet=rec)
    eager, graphed = make(), make()
    before = {k: v.clone() for k, v in graphed.recnet.state_dict().items()}
    graphed.capture_step(img1, img2, label, warmup=2)
    for k, v in graphed.recnet.state_dict().items():
        assert torch.equal(v, before[k]), "capture changed " + k
    assert graphed.optim.device_step() == 0
    for it in range(3):
        eager.step(img1, img2, label)
        graphed.step(img1, img2, label)
    torch.cuda.synchronize()
    for (k, a), (_, b) in zip(graphed.recnet.state_dict().items(), eager.recnet.state_dict().items()):
        assert torch.equal(a, b), k
    assert graphed.optim.device_step() == eager.optim.device_step() == 3
    assert int(graphed.recnet.state_dict()["Conv4Merge.0.norm.norm.num_batches_tracked"]) == 6
    # eval forward between replays
    y = torch.randn(3, 512, 7, 7, generator=torch.Generator().manual_seed(1)).cuda()
    graphed.recnet.eval()
    with torch.no_grad():
        v1, _ = graphed.recnet(y)
    graphed.recnet.train()
    for it in range(2):
        graphed.step(img1, img2, label)
    graphed.recnet.eval()
    fresh = RecNet()
    fresh.load_state_dict(graphed.recnet.state_dict())
    fresh = fresh.cuda().eval()
    with torch.no_grad():
        v2, _ = graphed.recnet(y)
        v3, _ = fresh(y)
    torch.cuda.synchronize()
    rel = lambda a, b: ((a - b).abs().max() / b.abs().max()).item()
    assert torch.equal(v2, v3)            # the eval path has no atomics either
    assert rel(v1, v2) >= 1e-3            # the two training steps in between did change the weights


@pytest.mark.parametrize("n", [4, 32])
def test_training_step_bit_reproducible_across_trainer_instances(lib, n):
    """Images -> frozen backbone -> RecNet forward / losses / backward: two Trainer instances (different buffers) and a
    repeated run produce bit-identical losses and gradients. (Before the backbone's SE squeeze lost its atomics, bf16 ulp
    differences in the feature map moved parameter gradients by up to 10 %: tools/trainer_instance_repro.py.)"""
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    a, b = ob.synth_faces(n, seed=100).cuda(), ob.synth_faces(n, seed=100, masked=True).cuda()
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
        return {k: p.grad.clone() for k, p in tr.recnet.named_parameters()}, [float(v) for v in tr.loss_items]
    t1, t2 = make(), make()
    g1, l1 = run(t1)
    g1b, l1b = run(t1)
    g2, l2 = run(t2)
    assert l1 == l1b == l2
    for k in g1:
        assert torch.equal(g1[k], g1b[k]), "run-to-run " + k
        assert torch.equal(g1[k], g2[k]), "across instances " + k


def test_trainer_full_step_public_api(lib):
    """The whole step through the public Trainer API (bf16 backbone in the loop), then clip + Adam + LR step."""
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    n = 4
    img1, img2 = ob.synth_faces(n, seed=5), ob.synth_faces(n, seed=5, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(5))
    items_ref, _, _, _ = otr.train_step(bsd, rsd, img1, img2, label)
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(), recnet=rec, encoder_weights=bsd)
    tr.step(img1.cuda(), img2.cuda(), label.cuda())
    vals = tr.get_current_values()
    print("trainer step:", vals, "oracle", items_ref)
    assert all(torch.isfinite(p).all() for p in rec.parameters())
    assert abs(float(vals["ClassifierLoss"]) - items_ref[3]) <= 2e-2 * items_ref[3]
    assert abs(float(vals["SelfSimilarityLoss"]) - items_ref[0]) <= 2e-2 * max(items_ref[0], 1e-3)
    clone = tr.clone_model()                                          # reference trainer.py:97-113
    assert all(torch.equal(a, b) for a, b in zip(clone["Recnet"].state_dict().values(), rec.state_dict().values()))


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


