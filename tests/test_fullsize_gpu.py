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
    # oracle on a slice (rows 448..455 = copies of faces 0..7)
    with torch.no_grad():
        y_ref, f_ref = ob.backbone_forward(bsd, base[:8])
        v_ref, _ = orr.recnet_forward(rsd, y_ref)
    rel = lambda a, b: ((a.cpu() - b).abs().max() / b.abs().max()).item()
    assert rel(f[448:456], f_ref) <= 1e-2 and rel(v[448:456], v_ref) <= 1e-2


def test_training_step_256_pairs_tiling_property(lib, nets):
    """Under batch-statistics BatchNorm a batch made of 4 copies of 64 pairs has the same statistics, the same mean
    losses and the same gradients as the 64 pairs alone. 256 pairs run on pixel-major tiles, 64 on row-major ones, so
    this also cross-checks the two tilings (forward, dgrad with tap skipping, wgrad) at the benchmark size."""
    from ffr_net_b200.trainer import Trainer, default_opts
    bsd, rsd, _, _ = nets
    assert lib.ffr_pixmajor_profitable(256) == 1 and lib.ffr_pixmajor_profitable(64) == 0
    a, b = ob.synth_faces(64, seed=41).cuda(), ob.synth_faces(64, seed=41, masked=True).cuda()
    label = torch.randint(0, 10575, (64,), generator=torch.Generator().manual_seed(41)).cuda()
    res = []
    for reps in (1, 4, 1):
        rec = RecNet()
        rec.load_state_dict(rsd)
        tr = Trainer(default_opts(lr=1e-3), recnet=rec, encoder_weights=bsd)
        tr.set_input(a.repeat(reps, 1, 1, 1), b.repeat(reps, 1, 1, 1), label.repeat(reps))
        tr.forward()
        tr.zero_grad()
        tr.backward()
        torch.cuda.synchronize()
        res.append(([float(l.detach()) for l in tr.loss_items], {k: p.grad.clone() for k, p in rec.named_parameters()},
                    float(tr._correct) / (64 * reps)))
    (l64, g64, acc64), (l256, g256, acc256), (l64b, g64b, _) = res
    assert all(torch.isfinite(g).all() for g in g256.values())
    assert all(abs(x - y) <= 2e-3 * max(1.0, abs(x)) for x, y in zip(l64, l256)), (l64, l256)
    assert abs(acc64 - acc256) <= 2.0 / 64

    def cosines(ga, gb):
        return sorted(torch.nn.functional.cosine_similarity(ga[k].reshape(1, -1), gb[k].reshape(1, -1)).item() for k in ga)
    c_tiled, c_noise = cosines(g256, g64), cosines(g64b, g64)
    print("grad cosine 256-tiled vs 64: min %.4f median %.4f | run-to-run at 64: min %.4f median %.4f" %
          (c_tiled[0], c_tiled[len(c_tiled) // 2], c_noise[0], c_noise[len(c_noise) // 2]))
    # gradients agree up to the bf16 rounding-noise level of the step itself
    assert c_tiled[len(c_tiled) // 2] >= 0.99 and c_tiled[0] >= min(0.9, c_noise[0] - 0.05)
