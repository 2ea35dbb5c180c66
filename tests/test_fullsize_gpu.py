"""GPU tests at BASELINE.json's full sizes (512 images per batch, 256 pairs per training step), where the CPU oracle is
too slow to run in full: size-independent properties (duplicate images, batch tiling under batch-statistics BatchNorm,
unit norms) plus an oracle check on a slice. These sizes take the large-batch code paths (pixel-major RecNet tiles,
multi-wave persistent grids, split heuristics) that the small parity cases do not reach."""
import pytest
import torch

from oracle import backbone as ob
from oracle import recnet as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nets(lib):
    bsd, rsd = ob.synth_backbone_state_dict(0), orr.synth_recnet_state_dict(0)
    enc = Backbone(50, 0.6, "ir_se")
    enc.load_state_dict(bsd)
    rec = RecNet()
    rec.load_state_dict(rsd)
    return bsd, rsd, enc.cuda().eval(), rec.cuda().eval()


def test_embeddings_bs512_properties(lib, nets):
    bsd, rsd, enc, rec = nets
    assert lib.ffr_pixmajor_profitable(512) == 1
    base = ob.synth_faces(64, seed=31)
    x = base.repeat(8, 1, 1, 1).cuda()                     # 512 images = 8 copies of 64 distinct faces
    with torch.no_grad():
        y, f = enc(x)
        v, fmap = rec(y)
        v2 = rec.embed_from_images(enc, x)
    torch.cuda.synchronize()
    assert all(torch.isfinite(t).all() for t in (y, f, v, fmap, v2))
    assert (f.norm(dim=1) - 1).abs().max().item() <= 1e-5                      # l2_norm, model_ir_se50.py:13-16
    # copies of one image agree wherever they sit in the batch: the bounds of the batch-invariance tests, with the
    # feature-map bound doubled because this is the extreme over 12.8 M elements (bf16 ulp flips through 24 units)
    yv, fv, vv = y.view(8, 64, -1), f.view(8, 64, -1), v.view(8, 64, -1)
    assert (yv - yv[:1]).abs().max().item() <= 2e-2 * y.abs().max().item()
    assert (fv - fv[:1]).abs().max().item() <= 2e-3
    assert (vv - vv[:1]).abs().max().item() <= 1e-2 * v.abs().max().item()
    assert (v2 - v).abs().max().item() <= 1e-2 * v.abs().max().item()
    cos = torch.nn.functional.cosine_similarity(vv[0], vv[5], dim=1)           # same image, different batch slot
    assert cos.min().item() >= 1 - 1e-3
    # fp32 CPU oracle on ALL 64 distinct faces of the batch, against their LAST copies (rows 448..511: the tail of every
    # persistent grid, the last image blocks of the pixel-major RecNet tiles)
    with torch.no_grad():
        y_ref, f_ref = ob.backbone_forward(bsd, base)
        v_ref, _ = orr.recnet_forward(rsd, y_ref)
    rel = lambda a, b: ((a.cpu() - b).abs().max() / b.abs().max()).item()
    e_y, e_f, e_v = rel(y[448:], y_ref), rel(f[448:], f_ref), rel(v[448:], v_ref)
    print("bs512, 64 distinct faces vs fp32 oracle: feature map %.2e embedding %.2e rectified %.2e" % (e_y, e_f, e_v))
    assert e_y <= 2e-2 and e_f <= 1e-2 and e_v <= 1e-2
    cos_f = torch.nn.functional.cosine_similarity(f[448:].cpu(), f_ref, dim=1)
    assert cos_f.min().item() >= 1 - 1e-3                                       # north star: embedding cosine within 1e-3


def test_training_step_256_pairs_tiling_property(lib, nets):
    """Under batch-statistics BatchNorm a batch made of 4 copies of 64 pairs has the same statistics, the same mean
    losses and the same gradients as the 64 pairs alone. The 256-pair step (BASELINE configs[2]) runs on pixel-major
    tiles, the 64-pair step is forced onto row-major tiles, so this also cross-checks the two tilings (forward, dgrad with
    tap skipping, wgrad) at the benchmark size; and the 256-pair step is reproducible bit for bit."""
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd, enc, _ = nets
    assert lib.ffr_pixmajor_profitable(512) == 1
    a, b = ob.synth_faces(64, seed=41).cuda(), ob.synth_faces(64, seed=41, masked=True).cuda()
    label = torch.randint(0, 10575, (64,), generator=torch.Generator().manual_seed(41)).cuda()
    with torch.no_grad():
        y64, f64 = enc(torch.cat((a, b)))
    res = []
    for reps, mode in ((1, 0), (4, -1), (4, -1)):
        lib.ffr_debug_set_pixmajor(mode)
        rec = RecNet()
        rec.load_state_dict(rsd)
        tr = Trainer(default_opts(lr=1e-3), recnet=rec, encoder_weights=bsd)
        yy = torch.cat((y64[:64].repeat(reps, 1, 1, 1), y64[64:].repeat(reps, 1, 1, 1)))
        ff = torch.cat((f64[:64].repeat(reps, 1), f64[64:].repeat(reps, 1)))
        tr.encoder = lambda x, yy=yy, ff=ff: (yy, ff)
        tr.set_input(a.repeat(reps, 1, 1, 1), b.repeat(reps, 1, 1, 1), label.repeat(reps))
        tr.forward()
        tr.backward()
        torch.cuda.synchronize()
        res.append(([float(l) for l in tr.loss_items], {k: p.grad.clone() for k, p in rec.named_parameters()},
                    float(tr._correct) / (64 * reps)))
    lib.ffr_debug_set_pixmajor(-1)
    (l64, g64, acc64), (l256, g256, acc256), (l256b, g256b, _) = res
    assert all(torch.isfinite(g).all() for g in g256.values())
    print("losses 64 pairs", l64, "256 pairs", l256)
    assert all(abs(x - y) <= 1e-4 * max(1.0, abs(x)) for x, y in zip(l64, l256)), (l64, l256)
    assert abs(acc64 - acc256) <= 1e-9
    errs = sorted(((g256[k].double() - g64[k].double()).norm() / (g64[k].double().norm() + 1e-30)).item() for k in g64)
    print("256-pair (4 x 64, pixel-major) vs 64-pair (row-major) gradients: worst %.3e median %.3e" % (errs[-1], errs[len(errs) // 2]))
    assert errs[-1] <= 4e-2 and errs[len(errs) // 2] <= 1e-2     # measured 1.7e-2 / 3.8e-3
    assert l256 == l256b and all(torch.equal(g256[k], g256b[k]) for k in g256)


def test_training_step_256_pairs_vs_oracle(lib, nets):
    """BASELINE configs[2] at full size against the PURE fp32 CPU oracle (oracle/train.py: both RecNet calls, the four
    losses, autograd): 256 distinct (unmasked, masked) pairs, RecNet fed with the oracle's backbone outputs so that only
    the training path is under test. Bounds = the 32-pair bounds of tests/test_train_gpu.py (5e-2 worst / 2e-2 median
    relative L2 per gradient tensor; the fp16 weight rounding of the forward GEMMs is the largest term)."""
    from oracle import train as otr
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd, enc, _ = nets
    n = 256
    img1, img2 = ob.synth_faces(n, seed=51), ob.synth_faces(n, seed=51, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(51))
    items_ref, grads_ref, stats_ref, acc_ref = otr.train_step(bsd, rsd, img1, img2, label)
    rec = RecNet()
    rec.load_state_dict(rsd)
    tr = Trainer(default_opts(lr=1e-3), recnet=rec, encoder_weights=bsd)
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
    items = [float(v) for v in tr.loss_items]
    print("256 pairs: losses", items, "fp32 oracle", items_ref)
    for a, b in zip(items, items_ref):
        assert abs(a - b) <= 1e-3 * max(abs(b), 1e-3)
    rl2 = lambda a, b: ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
    named = dict(rec.named_parameters())
    e = sorted(((rl2(p.grad.cpu(), grads_ref[k]), k) for k, p in named.items()), reverse=True)
    print("256 pairs vs pure fp32 oracle: worst %.3e (%s) median %.3e" % (e[0][0], e[0][1], e[len(e) // 2][0]))
    assert e[0][0] <= 5e-2 and e[len(e) // 2][0] <= 2e-2, e[:3]
    for k in ("Conv4Merge.0.norm.norm.running_mean", "Conv4Space.0.norm.norm.running_var"):
        assert rl2(rec.state_dict()[k].cpu(), stats_ref[k]) <= 2e-3, k
    assert abs(float(tr._correct) / n - acc_ref) <= 1.0 / n
