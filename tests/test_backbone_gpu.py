"""GPU parity of ffr_net_b200.Backbone (CUDA path through the C ABI) against the fp32 CPU oracle.

Tolerance (BASELINE.json north_star): embeddings <= 1e-2 max relative error in bf16 against fp32, measured as
max|e - e_ref| / max|e_ref| over the batch; pair cosine <= 1e-3 absolute."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import backbone as ob
from ffr_net_b200 import _lib
from ffr_net_b200.backbone import Backbone

pytestmark = pytest.mark.gpu

EMB_TOL = 1e-2
COS_TOL = 1e-3


@pytest.fixture(scope="module")
def models(lib):
    sd = ob.synth_backbone_state_dict(0)
    m = Backbone(50, 0.6, "ir_se")
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    return sd, m


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("n", [1, 5, 8])
def test_backbone_matches_oracle(models, n):
    sd, m = models
    x = ob.synth_faces(n, seed=n)
    with torch.no_grad():
        y_ref, f_ref = ob.backbone_forward(sd, x)
        y, f = m(x.cuda())
    torch.cuda.synchronize()
    y, f = y.cpu(), f.cpu()
    assert y.shape == (n, 512, 7, 7) and f.shape == (n, 512)
    assert torch.isfinite(y).all() and torch.isfinite(f).all()
    print("n=%d embedding rel err %.3e  featmap rel err %.3e" % (n, _rel(f, f_ref), _rel(y, y_ref)))
    assert _rel(f, f_ref) <= EMB_TOL
    assert _rel(y, y_ref) <= 1.5e-2          # feature map: looser, it is an intermediate (bf16 residual stream)
    assert (f.norm(dim=1) - 1).abs().max().item() <= 1e-5
    assert (F.cosine_similarity(f, f_ref) - 1).abs().max().item() <= COS_TOL


def test_pair_cosine_and_masked_inputs(models):
    """Cosine scores of (clean, masked) pairs agree with the oracle within 1e-3 absolute."""
    sd, m = models
    a = ob.synth_faces(6, seed=3)
    b = ob.synth_faces(6, seed=3, masked=True)
    with torch.no_grad():
        _, fa_ref = ob.backbone_forward(sd, a)
        _, fb_ref = ob.backbone_forward(sd, b)
        _, fa = m(a.cuda())
        _, fb = m(b.cuda())
    cos_ref = F.cosine_similarity(fa_ref, fb_ref)
    cos = F.cosine_similarity(fa.cpu(), fb.cpu())
    assert (cos - cos_ref).abs().max().item() <= COS_TOL


def test_backbone_batch_invariance(models):
    """Images are independent: row i of a batch equals the same image run alone, up to fp32 accumulation order (the SE
    squeeze sums 32-row blocks whose boundaries depend on where the image sits in the batch, the head's K splits depend on
    the batch size): the feature map may differ by a bf16 ulp here and there that then propagates through the remaining
    units (measured up to 6e-3 of its range here, up to 1e-2 over larger batches), the unit-norm embedding by a few
    1e-4. The SAME batch run twice is bit-identical (test_backbone_bit_reproducible)."""
    sd, m = models
    x = ob.synth_faces(4, seed=9).cuda()
    with torch.no_grad():
        y4, f4 = m(x)
        y1, f1 = m(x[2:3])
    dy = (y4[2:3] - y1).abs().max().item() / y1.abs().max().item()
    df = (f4[2:3] - f1).abs().max().item()
    print("batch invariance: featmap rel diff %.3e, embedding abs diff %.3e" % (dy, df))
    assert dy <= 2e-2 and df <= 2e-3     # chaotic amplification of single bf16 ulps through 24 units (measured <= 5.8e-3 / 3.4e-4)


@pytest.mark.parametrize("n", [1, 37, 160])
def test_backbone_bit_reproducible(models, n):
    """No atomics on the eval path: the SE squeeze is stored as per-32-row-block partial sums and added per image in a
    fixed order, the head's split-K partial products are added in split order. The same batch gives the same bits run
    after run, from a second Backbone instance (other buffers), and on a side stream under load."""
    from ffr_net_b200.backbone import Backbone
    sd, m = models
    x = ob.synth_faces(n, seed=5).cuda()
    other = Backbone(50, 0.6, "ir_se")
    other.load_state_dict(sd)
    other = other.cuda().eval()
    with torch.no_grad():
        y0, f0 = m(x)
        y0, f0 = y0.clone(), f0.clone()
        for _ in range(3):
            y1, f1 = m(x)
            assert torch.equal(y1, y0) and torch.equal(f1, f0)
        y2, f2 = other(x)
        assert torch.equal(y2, y0) and torch.equal(f2, f0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        busy = torch.randn(4096, 4096, device="cuda")
        with torch.cuda.stream(side):
            y3, f3 = other(x)
        for _ in range(4):
            busy = busy @ busy * 1e-3            # concurrent work on the main stream changes CTA scheduling
        torch.cuda.synchronize()
        assert torch.equal(y3, y0) and torch.equal(f3, f0)


def test_backbone_rejects_training_and_cpu(models):
    sd, m = models
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 112, 112))
    m.train()
    try:
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 3, 112, 112, device="cuda"))
    finally:
        m.eval()


@pytest.mark.parametrize("n,S", [(2, 112), (3, 14)])
def test_stem_kernel(lib, n, S):
    """ffr_stem_fwd (warp-MMA im2col) vs F.conv2d on bf16-rounded operands; tolerance = bf16 output rounding."""
    from ffr_net_b200 import _lib, layout
    g = torch.Generator().manual_seed(S)
    x = torch.randn(n, 3, S, S, generator=g).clamp_(-1, 1).cuda()
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).cuda()
    b = torch.randn(64, generator=g).cuda() * 0.1
    a = torch.empty(64).uniform_(0.1, 0.4, generator=g).cuda()
    wk = w.reshape(64, 27).t().contiguous()
    out = torch.full((n * (S + 1) * (S + 1), 64), 9.0, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.ffr_stem_fwd(_lib.ptr(x), _lib.ptr(wk), _lib.ptr(b), _lib.ptr(a), _lib.ptr(out), n, S,
                                _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), padding=1) + b.view(1, -1, 1, 1)
    ref = torch.where(ref > 0, ref, ref * a.view(1, -1, 1, 1))
    got = layout.from_flat(out, n, S, 64)
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    assert layout.flat_pad_rows(out, n, S, 64).abs().max().item() == 0.0


@pytest.mark.parametrize("n,S", [(3, 112), (2, 16), (1, 48)])
def test_stem_strip_kernel_matches_gather_kernel(lib, n, S):
    """The shared-memory strip kernel (S % 16 == 0) and the gather kernel (any S) are the same arithmetic: bit-identical
    maps, for fp32 NCHW input and for uint8 HWC input with flips and channel swap; pad rows are zero."""
    from oracle import preprocess as opp
    from ffr_net_b200 import _lib, layout
    g = torch.Generator().manual_seed(100 + S)
    x = torch.randn(n, 3, S, S, generator=g).clamp_(-1, 1).cuda()
    w = (torch.randn(64, 3, 3, 3, generator=g) * 0.2).reshape(64, 27).t().contiguous().cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    a = torch.empty(64).uniform_(0.1, 0.4, generator=g).cuda()
    imgs = torch.from_numpy(opp.synth_images_u8(n, S, seed=3)).cuda()
    flips = torch.tensor([1, 0, 1][:n], dtype=torch.uint8).cuda()
    rows = n * (S + 1) * (S + 1)
    st = _lib.stream_ptr()
    outs = {}
    try:
        for strip in (1, 0):
            lib.ffr_debug_set_stem_strip(strip)
            o = torch.full((rows, 64), 9.0, dtype=torch.bfloat16, device="cuda")
            u = torch.full((rows, 64), 9.0, dtype=torch.bfloat16, device="cuda")
            _lib.check(lib.ffr_stem_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(a), _lib.ptr(o), n, S, st))
            _lib.check(lib.ffr_stem_u8_fwd(_lib.ptr(imgs), _lib.ptr(flips), 1, _lib.ptr(w), _lib.ptr(b), _lib.ptr(a),
                                           _lib.ptr(u), n, S, st))
            torch.cuda.synchronize()
            outs[strip] = (o, u)
    finally:
        lib.ffr_debug_set_stem_strip(1)
    assert torch.equal(outs[1][0], outs[0][0]) and torch.equal(outs[1][1], outs[0][1])
    assert layout.flat_pad_rows(outs[1][0], n, S, 64).abs().max().item() == 0.0
    assert layout.from_flat(outs[1][0], n, S, 64).abs().max().item() > 0.1


def test_empty_and_odd_batches(models):
    """Edge cases: empty batch, batch sizes that leave ragged last tiles (rows not a multiple of 128/256)."""
    sd, m = models
    with torch.no_grad():
        y, f = m(torch.zeros(0, 3, 112, 112, device="cuda"))
        assert y.shape == (0, 512, 7, 7) and f.shape == (0, 512)
        x = ob.synth_faces(7, seed=4)
        y7, f7 = m(x.cuda())
        y3, f3 = m(x[:3].cuda())
    assert torch.isfinite(f7).all()
    assert (f7[:3] - f3).abs().max().item() <= 2e-3          # batch-invariance bound
    with pytest.raises(ValueError):
        m(torch.zeros(2, 3, 96, 112, device="cuda"))


def test_stem_u8_fused_preprocessing(lib):
    """ffr_stem_u8_fwd (decoded uint8 HWC images, channel swap + per-image flip + ToTensor/Normalize fused) produces
    bit-identical stem activations to ffr_stem_fwd on the oracle-preprocessed fp32 NCHW tensor."""
    from oracle import preprocess as opp
    from ffr_net_b200 import packing
    sd = ob.synth_backbone_state_dict(0)
    from ffr_net_b200.backbone import Backbone, _bn_fold
    m = Backbone(50, 0.6, "ir_se")
    m.load_state_dict(sd)
    conv, bn, prelu = m.input_layer[0], m.input_layer[1], m.input_layer[2]
    w, b = packing.pack_stem(conv.weight.detach(), _bn_fold(bn))
    w, b, a = w.cuda(), b.cuda(), prelu.weight.detach().float().cuda()
    for n, S in ((3, 16), (2, 112)):
        imgs = opp.synth_images_u8(n, S, seed=9)
        flips = np.array([1, 0, 1][:n], dtype=np.uint8)
        for swap in (1, 0):
            x = opp.preprocess_batch(imgs, flips, swap_rb=bool(swap)).cuda()
            rows = n * (S + 1) * (S + 1)
            o0 = torch.empty(rows, 64, dtype=torch.bfloat16, device="cuda")
            o1 = torch.empty(rows, 64, dtype=torch.bfloat16, device="cuda")
            st = _lib.stream_ptr()
            _lib.check(lib.ffr_stem_fwd(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(a), _lib.ptr(o0), n, S, st))
            iu, fl = torch.from_numpy(imgs).cuda(), torch.from_numpy(flips).cuda()
            _lib.check(lib.ffr_stem_u8_fwd(_lib.ptr(iu), _lib.ptr(fl), swap, _lib.ptr(w), _lib.ptr(b), _lib.ptr(a),
                                           _lib.ptr(o1), n, S, st))
            torch.cuda.synchronize()
            assert torch.equal(o0, o1)


def test_backbone_forward_u8_matches_fp32_input(models):
    """Backbone(uint8 NHWC) == Backbone(preprocess(uint8)) (no flip, channel swap as data/dataset.py:138-141)."""
    from oracle import preprocess as opp
    sd, m = models
    imgs = opp.synth_images_u8(3, 112, seed=4)
    x = opp.preprocess_batch(imgs).cuda()
    with torch.no_grad():
        y0, f0 = m(x)
        y1, f1 = m(torch.from_numpy(imgs).cuda())
        flips = torch.tensor([1, 0, 1], dtype=torch.uint8)
        y2, f2 = m.forward_u8(torch.from_numpy(imgs).cuda(), flip=flips.cuda())
        y3, f3 = m(opp.preprocess_batch(imgs, flips.numpy()).cuda())
    # same kernels on bit-identical stem inputs: only the fp32 atomics order of the SE / head sums differs
    assert (f1 - f0).abs().max().item() <= 2e-3 and (y1 - y0).abs().max().item() <= 2e-2 * y0.abs().max().item()
    assert (f2 - f3).abs().max().item() <= 2e-3 and (y2 - y3).abs().max().item() <= 2e-2 * y3.abs().max().item()
    assert (f2[1] - f0[1]).abs().max().item() <= 2e-3          # image 1 is not flipped
    assert (f2[0] - f0[0]).abs().max().item() > 1e-2            # image 0 is


def test_launch_modes_bit_identical(models, lib):
    """Programmatic dependent launch (a kernel's CTAs are scheduled while its predecessor drains) and CTA pairs
    (cta_group::2) change scheduling, not arithmetic, and the backbone-only ("lean") epilogue instantiation is the same
    arithmetic with the unused features compiled out: the forward is bit-identical with any of them switched off, and
    repeated back-to-back forwards (no host sync in between: the dependent-launch race window) stay identical."""
    _, m = models
    x = ob.synth_faces(16, seed=21).repeat(4, 1, 1, 1).cuda()      # 64 images: several tiles per layer
    with torch.no_grad():
        y0, f0 = m(x)
        outs = [m(x) for _ in range(4)]                                # back to back, same workspace
        torch.cuda.synchronize()
        for y, f in outs:
            assert torch.equal(y, y0) and torch.equal(f, f0)
        try:
            lib.ffr_debug_set_pdl(0)
            y1, f1 = m(x)
            lib.ffr_debug_set_pdl(3)
            lib.ffr_debug_set_lean_epilogue(0)
            y3, f3 = m(x)
            lib.ffr_debug_set_lean_epilogue(1)
            lib.ffr_debug_set_pair(0)
            y2, f2 = m(x)
            lib.ffr_debug_set_pair(-1)
            # opt-in stream-K schedule: split work items are finished as (own + peer) partial sums, i.e. another fp32
            # summation order than whole items: equal within the bf16 noise, and bit-reproducible in itself
            lib.ffr_debug_set_streamk(1)
            y4, f4 = m(x)
            y5, f5 = m(x)
        finally:
            lib.ffr_debug_set_pdl(-1)
            lib.ffr_debug_set_pair(-1)
            lib.ffr_debug_set_lean_epilogue(1)
            lib.ffr_debug_set_streamk(0)
        torch.cuda.synchronize()
    assert torch.equal(y1, y0) and torch.equal(f1, f0)
    assert torch.equal(y3, y0) and torch.equal(f3, f0)
    assert torch.equal(y2, y0) and torch.equal(f2, f0)
    assert torch.equal(y5, y4) and torch.equal(f5, f4)
    assert (f4 - f0).abs().max().item() <= 2e-3 and (y4 - y0).abs().max().item() <= 2e-2 * y0.abs().max().item()
