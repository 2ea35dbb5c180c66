"""GPU parity of the tcgen05 implicit-GEMM kernel (through the C ABI) against torch fp32 on bf16-rounded operands.

Integer-exact checks where the arithmetic allows (small-integer operands make fp32 accumulation exact), tolerance
2e-3 relative otherwise (fp32 accumulation-order differences on bf16 operands; stated per test)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from ffr_net_b200 import _lib, layout
import emulate

pytestmark = pytest.mark.gpu

P = _lib.ptr
EPI = dict(BIAS=1, BORDER=2, PRELU=4, GEOM=8, POOL=16, S2D=32, F32ATOMIC=64, SIGMOID=128, SCATTER=256, RESIDUAL=512,
           STATS=1024, F32=2048)


def _stream():
    return _lib.stream_ptr()


@pytest.fixture(params=["window", "window1", "tiles"], autouse=True)
def conv_variant(request, lib):
    """Every test runs with the sliding-window kernels enabled (CTA pairs on the 256-wide N tiles), with the
    single-CTA window kernel everywhere ("window1"), and with the tile-per-tap kernel forced."""
    lib.ffr_debug_set_window(0 if request.param == "tiles" else 1)
    lib.ffr_debug_set_pair(0 if request.param == "window1" else -1)
    yield request.param
    lib.ffr_debug_set_window(1)
    lib.ffr_debug_set_pair(-1)


def _gemm(lib, a, wp, cin, cout, taps, M, flags=0, bias=None, slope=None, out=None, ldo=0, geom=(64, 1, 1, 0, 1),
          s2d_so=0, pool=None, out_f32=None, res=None, ldres=0, stats=None, splits=1, scatter=None, scatter_n=0,
          out_rpi=0, b_rows_per_mtile=0):
    shifts = (ctypes.c_int * 9)(*([t[0] for t in taps] + [0] * (9 - len(taps))))
    choffs = (ctypes.c_int * 9)(*([t[1] for t in taps] + [0] * (9 - len(taps))))
    rpi, wp_, s, h0, nimg = geom
    rc = lib.ffr_conv_gemm(P(a), a.shape[0], a.shape[1], a.stride(0), P(wp), cin, cout, len(taps), shifts, choffs, M,
                           rpi, wp_, s, h0, nimg, flags, P(bias), P(slope), P(out), ldo, s2d_so, P(pool), P(out_f32),
                           P(res), ldres, P(stats), splits, P(scatter), scatter_n, out_rpi, b_rows_per_mtile, _stream())
    _lib.check(rc, "ffr_conv_gemm")


@pytest.mark.parametrize("cout", [64, 128, 256, 512])
@pytest.mark.parametrize("M,K", [(128, 64), (300, 192), (1000, 576)])
def test_plain_gemm_exact(lib, cout, M, K):
    """Small-integer operands: every product and partial sum is exactly representable -> bit-exact vs torch."""
    g = torch.Generator(device="cuda").manual_seed(M + K + cout)
    a = torch.randint(-3, 4, (M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randint(-2, 3, (cout, K), generator=g, device="cuda").to(torch.bfloat16)
    out = torch.full((M, cout), 7.0, dtype=torch.bfloat16, device="cuda")
    _gemm(lib, a, w, K, cout, [(0, 0)], M, out=out, ldo=cout)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t())
    assert torch.equal(out.float(), ref.to(torch.bfloat16).float())


def test_gemm_f32_out_and_bias(lib):
    M, K, cout = 515, 256, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, K, generator=g, device="cuda") * 0.1).to(torch.bfloat16)
    bias = torch.randn(cout, generator=g, device="cuda")
    out = torch.zeros(M, cout, dtype=torch.float32, device="cuda")
    _gemm(lib, a, w, K, cout, [(0, 0)], M, flags=EPI["BIAS"] | EPI["F32"], bias=bias, out_f32=out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    assert (out - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("splits", [2, 5, 16])
def test_splitk_atomic(lib, splits):
    M, K, cout = 200, 2048, 512
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randint(-2, 3, (M, K), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randint(-2, 3, (cout, K), generator=g, device="cuda").to(torch.bfloat16)
    out = torch.zeros(M, cout, dtype=torch.float32, device="cuda")
    _gemm(lib, a, w, K, cout, [(0, 0)], M, flags=EPI["F32ATOMIC"], out_f32=out, splits=splits)
    torch.cuda.synchronize()
    assert torch.equal(out, a.float() @ w.float().t())   # integer-valued: exact in any order


@pytest.fixture(params=["window", "pixmajor"])
def backbone_tiles(request, lib):
    """Backbone 3x3/s1 convolutions on the production sliding-window tiles and on the experimental pixel-major tiles
    over the halo-shared flat layout (zero padding from TMA out-of-bounds fill, pad points stored as zero tiles)."""
    lib.ffr_debug_set_pixmajor_backbone(1000 if request.param == "pixmajor" else 0)
    yield request.param
    lib.ffr_debug_set_pixmajor_backbone(0)


@pytest.mark.parametrize("n,S,cin,cout", [(3, 14, 64, 64), (2, 7, 128, 256), (5, 28, 64, 128), (2, 14, 256, 512),
                                            (2, 112, 64, 64), (3, 56, 64, 128)])
def test_conv3x3_bnpre_prelu(lib, n, S, cin, cout, backbone_tiles):
    """conv(pad0(BN(x))) + PReLU through ffr_conv3x3_bnpre_prelu_fwd vs F.conv2d on the same bf16-rounded operands.
    Tolerance 2e-3 of max|ref| (fp32 accumulation order) + bf16 output rounding (2^-8 relative)."""
    from ffr_net_b200 import packing
    g = torch.Generator(device="cuda").manual_seed(n * S + cin)
    x = torch.randn(n, cin, S, S, generator=g, device="cuda")
    w = torch.randn(cout, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
    s0 = torch.empty(cin, device="cuda").uniform_(0.8, 1.2, generator=g)
    b0 = torch.empty(cin, device="cuda").uniform_(-0.3, 0.3, generator=g)
    slope = torch.empty(cout, device="cuda").uniform_(0.1, 0.4, generator=g)
    wp = packing.pack_conv(w, in_scale=s0)
    bias9 = packing.border_bias_table(w, b0)
    xf = layout.to_flat(x)
    out = torch.full((xf.shape[0], cout), 3.0, dtype=torch.bfloat16, device="cuda")
    rc = lib.ffr_conv3x3_bnpre_prelu_fwd(P(xf), n, S, cin, P(wp), cout, P(bias9), P(slope), P(out), 0, _stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    # reference on the operands the kernel sees: bf16 x, bf16 folded weights; the shift part in fp32
    xb = layout.from_flat(xf, n, S, cin)
    wb = wp.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2)
    shift_map = b0.view(1, -1, 1, 1).expand(n, cin, S, S)
    ref = F.conv2d(xb, wb, padding=1) + F.conv2d(shift_map, w, padding=1)
    ref = torch.where(ref > 0, ref, ref * slope.view(1, -1, 1, 1))
    got = layout.from_flat(out, n, S, cout)
    tol = 2e-3 * ref.abs().max().item() + 2 ** -8 * ref.abs().max().item()
    assert (got - ref).abs().max().item() <= tol
    assert layout.flat_pad_rows(out, n, S, cout).abs().max().item() == 0.0


@pytest.mark.parametrize("n,S,c,cout,stride", [(3, 14, 64, 64, 1), (2, 14, 128, 128, 2), (4, 28, 64, 64, 2),
                                                (2, 7, 512, 512, 1)])
def test_conv3x3_bn_pool(lib, n, S, c, cout, stride, backbone_tiles):
    from ffr_net_b200 import packing
    g = torch.Generator(device="cuda").manual_seed(S + c + stride)
    x = torch.randn(n, c, S, S, generator=g, device="cuda")
    w = torch.randn(cout, c, 3, 3, generator=g, device="cuda") / (3 * c ** 0.5)
    s1 = torch.empty(cout, device="cuda").uniform_(0.8, 1.2, generator=g)
    b1 = torch.empty(cout, device="cuda").uniform_(-0.3, 0.3, generator=g)
    wp = packing.pack_conv(w, out_scale=s1)
    xin = layout.to_flat(x) if stride == 1 else layout.to_s2d(x)
    so = S // stride
    rows = n * (so + 1) * (so + 1)
    out = torch.full((rows, cout), 3.0, dtype=torch.bfloat16, device="cuda")
    # squeeze partial sums: never zeroed by the caller (garbage-filled here); ffr_se_gate_fwd adds them per image
    part = torch.full((max(lib.ffr_se_pool_part_floats(n, so, cout), n * cout),), 7.0e3, dtype=torch.float32, device="cuda")
    rc = lib.ffr_conv3x3_bn_pool_fwd(P(xin), n, S, c, stride, P(wp), cout, P(b1), P(out), P(part), _stream())
    _lib.check(rc)
    fc1 = torch.randn(cout // 16, cout, generator=g, device="cuda") / cout ** 0.5
    fc2 = torch.randn(cout, cout // 16, generator=g, device="cuda") / 2
    gate = torch.empty(n, cout, device="cuda")
    pool = torch.empty(n, cout, device="cuda")
    _lib.check(lib.ffr_se_gate_fwd(P(part), P(fc1), P(fc2), P(gate), P(pool), n, so, cout, _stream()))
    torch.cuda.synchronize()
    xb = x.to(torch.bfloat16).float()
    wb = wp.float().reshape(cout, 3, 3, c).permute(0, 3, 1, 2)
    ref = F.conv2d(xb, wb, stride=stride, padding=1) + b1.view(1, -1, 1, 1)
    got = layout.from_flat(out, n, so, cout)
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= (2e-3 + 2 ** -8) * scale
    assert layout.flat_pad_rows(out, n, so, cout).abs().max().item() == 0.0
    ref_pool = ref.sum(dim=(2, 3))
    assert (pool - ref_pool).abs().max().item() <= 2e-3 * scale * so * so ** 0.5 + 1e-3
    # gate = sigmoid(W2 relu(W1 mean)) (model_ir_se50.py:29-36) on the kernel's own sums: fp32, 1e-5
    ref_gate = torch.sigmoid(torch.relu((pool / (so * so)) @ fc1.t()) @ fc2.t())
    assert (gate - ref_gate).abs().max().item() <= 1e-5
    if backbone_tiles != "pixmajor":       # row-major tiles: plain stores, fixed order -> bit-reproducible
        out2, pool2, gate2 = torch.empty_like(out), torch.empty_like(pool), torch.empty_like(gate)
        part.fill_(-3.0)
        _lib.check(lib.ffr_conv3x3_bn_pool_fwd(P(xin), n, S, c, stride, P(wp), cout, P(b1), P(out2), P(part), _stream()))
        _lib.check(lib.ffr_se_gate_fwd(P(part), P(fc1), P(fc2), P(gate2), P(pool2), n, so, cout, _stream()))
        torch.cuda.synchronize()
        assert torch.equal(out, out2) and torch.equal(pool, pool2) and torch.equal(gate, gate2)


@pytest.mark.parametrize("n,S,c,cout", [(345, 14, 256, 256), (330, 7, 512, 512), (512, 14, 256, 256)])
def test_conv3x3_stream_k(lib, n, S, c, cout, conv_variant):
    """Stream-K scheduling of the CTA-pair window kernel (ffr_set_conv_scratch): the (item, k-chunk) steps are split evenly
    over the pairs and split items are finished as own + peer partial. Against F.conv2d within the usual tolerance,
    within one bf16 rounding of the whole-item schedule, bit-reproducible, flag words left zero, squeeze sums right."""
    if conv_variant != "window":
        pytest.skip("CTA-pair window kernel only")
    from ffr_net_b200 import packing
    g = torch.Generator(device="cuda").manual_seed(S + c + n)
    x = torch.randn(n, c, S, S, generator=g, device="cuda")
    w = torch.randn(cout, c, 3, 3, generator=g, device="cuda") / (3 * c ** 0.5)
    s1 = torch.empty(cout, device="cuda").uniform_(0.8, 1.2, generator=g)
    b1 = torch.empty(cout, device="cuda").uniform_(-0.3, 0.3, generator=g)
    wp = packing.pack_conv(w, out_scale=s1)
    xin = layout.to_flat(x)
    rows = n * (S + 1) * (S + 1)
    scratch = torch.zeros(lib.ffr_conv_scratch_bytes(), dtype=torch.uint8, device="cuda")
    fc1 = torch.randn(cout // 16, cout, generator=g, device="cuda") / cout ** 0.5
    fc2 = torch.randn(cout, cout // 16, generator=g, device="cuda") / 2

    def run(use_scratch):
        out = torch.full((rows, cout), 3.0, dtype=torch.bfloat16, device="cuda")
        part = torch.full((max(lib.ffr_se_pool_part_floats(n, S, cout), n * cout),), 7.0e3, dtype=torch.float32, device="cuda")
        pool, gate = torch.empty(n, cout, device="cuda"), torch.empty(n, cout, device="cuda")
        _lib.check(lib.ffr_set_conv_scratch(P(scratch) if use_scratch else None, scratch.numel() if use_scratch else 0))
        lib.ffr_debug_set_streamk(1)              # opt-in schedule (default off)
        try:
            _lib.check(lib.ffr_conv3x3_bn_pool_fwd(P(xin), n, S, c, 1, P(wp), cout, P(b1), P(out), P(part), _stream()))
        finally:
            lib.ffr_set_conv_scratch(None, 0)
            lib.ffr_debug_set_streamk(0)
        assert lib.ffr_debug_last_streamk() == (1 if use_scratch else 0)      # the schedule under test really ran
        _lib.check(lib.ffr_se_gate_fwd(P(part), P(fc1), P(fc2), P(gate), P(pool), n, S, cout, _stream()))
        torch.cuda.synchronize()
        return out, pool

    o_sk, p_sk = run(True)
    assert int(scratch[:1024].max()) == 0                       # every flag consumed
    o_sk2, p_sk2 = run(True)
    o_it, p_it = run(False)
    assert torch.equal(o_sk, o_sk2) and torch.equal(p_sk, p_sk2)
    xb = x.to(torch.bfloat16).float()
    wb = wp.float().reshape(cout, 3, 3, c).permute(0, 3, 1, 2)
    ref = F.conv2d(xb, wb, padding=1) + b1.view(1, -1, 1, 1)
    scale = ref.abs().max().item()
    assert (layout.from_flat(o_sk, n, S, cout) - ref).abs().max().item() <= (2e-3 + 2 ** -8) * scale
    assert layout.flat_pad_rows(o_sk, n, S, cout).abs().max().item() == 0.0
    assert (o_sk.float() - o_it.float()).abs().max().item() <= 2 ** -7 * scale
    assert (p_sk - ref.sum(dim=(2, 3))).abs().max().item() <= 2e-3 * scale * S * S ** 0.5 + 1e-3


@pytest.mark.parametrize("n,S,C,mode", [(3, 14, 256, 0), (2, 7, 512, 0), (5, 28, 64, 1), (2, 56, 64, 2), (1, 7, 128, 2)])
def test_se_residual_isolated(lib, n, S, C, mode):
    """ffr_se_residual_fwd alone: y = u * gate[n] + shortcut on the flat map (model_ir_se50.py:36,73-76), against torch on
    the same bf16 inputs; bound: one bf16 rounding of the result. mode 1 reads the shortcut from the (2S)x(2S) grid
    (MaxPool2d(1, 2) = subsampling), mode 0 / 2 from the same grid."""
    g = torch.Generator(device="cuda").manual_seed(S * C + mode)
    u = torch.randn(n, C, S, S, generator=g, device="cuda")
    gate = torch.rand(n, C, generator=g, device="cuda")
    S_sc = 2 * S if mode == 1 else S
    sc = torch.randn(n, C, S_sc, S_sc, generator=g, device="cuda")
    uf, scf = layout.to_flat(u), layout.to_flat(sc)
    y = torch.full_like(uf, 9.0)
    _lib.check(lib.ffr_se_residual_fwd(P(uf), P(gate), P(scf), mode, P(y), n, S, C, _stream()))
    torch.cuda.synchronize()
    ub = layout.from_flat(uf, n, S, C)
    sb = layout.from_flat(scf, n, S_sc, C)
    if mode == 1:
        sb = sb[:, :, ::2, ::2]
    ref = ub * gate.view(n, C, 1, 1) + sb
    got = layout.from_flat(y, n, S, C)
    assert (got - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    assert layout.flat_pad_rows(y, n, S, C).abs().max().item() == 0.0          # pad rows: 0 * gate + 0


@pytest.mark.parametrize("n,S,C,mode", [(3, 14, 256, 0), (2, 7, 512, 0), (5, 28, 64, 1), (2, 56, 64, 2), (3, 28, 128, 0),
                                        (130, 14, 256, 0)])
def test_se_gate_residual_fused_matches_two_calls(lib, n, S, C, mode):
    """ffr_se_gate_residual_fwd (gate computed per image inside the streaming kernel) == ffr_se_gate_fwd followed by
    ffr_se_residual_fwd, bit for bit, on the squeeze partial sums a real conv2 launch stored."""
    from ffr_net_b200 import packing
    g = torch.Generator(device="cuda").manual_seed(S * C + mode + n)
    x = torch.randn(n, C, S, S, generator=g, device="cuda")
    w = torch.randn(C, C, 3, 3, generator=g, device="cuda") / (3 * C ** 0.5)
    wp = packing.pack_conv(w)
    b1 = torch.empty(C, device="cuda").uniform_(-0.3, 0.3, generator=g)
    rows = n * (S + 1) * (S + 1)
    u = torch.empty((rows, C), dtype=torch.bfloat16, device="cuda")
    part = torch.full((max(lib.ffr_se_pool_part_floats(n, S, C), n * C),), -5.0e3, dtype=torch.float32, device="cuda")
    _lib.check(lib.ffr_conv3x3_bn_pool_fwd(P(layout.to_flat(x)), n, S, C, 1, P(wp), C, P(b1), P(u), P(part), _stream()))
    fc1 = torch.randn(C // 16, C, generator=g, device="cuda") / C ** 0.5
    fc2 = torch.randn(C, C // 16, generator=g, device="cuda") / 2
    S_sc = 2 * S if mode == 1 else S
    scf = layout.to_flat(torch.randn(n, C, S_sc, S_sc, generator=g, device="cuda"))
    gate = torch.empty(n, C, device="cuda")
    y1, y2 = torch.full_like(u, 9.0), torch.full_like(u, -9.0)
    _lib.check(lib.ffr_se_gate_fwd(P(part), P(fc1), P(fc2), P(gate), None, n, S, C, _stream()))
    _lib.check(lib.ffr_se_residual_fwd(P(u), P(gate), P(scf), mode, P(y1), n, S, C, _stream()))
    _lib.check(lib.ffr_se_gate_residual_fwd(P(u), P(part), P(fc1), P(fc2), P(scf), mode, P(y2), n, S, C, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    assert y1.float().abs().max().item() > 0.1


def test_conv3x3_s2d_output_roundtrip(lib):
    """conv1 writing the space-to-depth layout, then the stride-2 conv reading it == conv -> prelu -> stride-2 conv."""
    from ffr_net_b200 import packing
    n, S, cin, c = 2, 28, 64, 128
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, cin, S, S, generator=g, device="cuda")
    w1 = torch.randn(c, cin, 3, 3, generator=g, device="cuda") / (3 * cin ** 0.5)
    w2 = torch.randn(c, c, 3, 3, generator=g, device="cuda") / (3 * c ** 0.5)
    slope = torch.empty(c, device="cuda").uniform_(0.1, 0.4, generator=g)
    zeros9 = torch.zeros(9, c, device="cuda")
    zb = torch.zeros(c, device="cuda")
    wp1, wp2 = packing.pack_conv(w1), packing.pack_conv(w2)
    so = S // 2
    t = torch.zeros(n * (so + 1) ** 2, 4 * c, dtype=torch.bfloat16, device="cuda")
    out = torch.empty(n * (so + 1) ** 2, c, dtype=torch.bfloat16, device="cuda")
    xf = layout.to_flat(x)
    _lib.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(xf), n, S, cin, P(wp1), c, P(zeros9), P(slope), P(t), 1, _stream()))
    _lib.check(lib.ffr_conv3x3_bn_pool_fwd(P(t), n, S, c, 2, P(wp2), c, P(zb), P(out), None, _stream()))
    torch.cuda.synchronize()
    xb = x.to(torch.bfloat16).float()
    r1 = F.conv2d(xb, wp1.float().reshape(c, 3, 3, cin).permute(0, 3, 1, 2), padding=1)
    r1 = torch.where(r1 > 0, r1, r1 * slope.view(1, -1, 1, 1)).to(torch.bfloat16).float()
    assert (layout.from_s2d(t, n, S, c) - r1).abs().max().item() <= 2 ** -7 * r1.abs().max().item()
    t_exact = layout.from_s2d(t, n, S, c)
    r2 = F.conv2d(t_exact, wp2.float().reshape(c, 3, 3, c).permute(0, 3, 1, 2), stride=2, padding=1)
    got = layout.from_flat(out, n, so, c)
    assert (got - r2).abs().max().item() <= (2e-3 + 2 ** -8) * r2.abs().max().item()


def test_emulation_matches_kernel(lib):
    """The torch emulation used by the CPU tests is the same function as the kernel (ties CPU tests to the device)."""
    n, S, cin, cout = 2, 14, 64, 64
    G = S + 1
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-2, 3, (n * G * G, cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randint(-2, 3, (cout, 9 * cin), generator=g, device="cuda").to(torch.bfloat16)
    out = torch.zeros(n * G * G, cout, dtype=torch.float32, device="cuda")
    taps = emulate.taps_3x3_flat(G)
    _gemm(lib, a, w, cin, cout, taps, n * G * G, flags=EPI["F32"], out_f32=out)
    torch.cuda.synchronize()
    assert torch.equal(out, emulate.conv_gemm(a, w, cin, taps))
