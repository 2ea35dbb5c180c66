"""Per-kernel GPU parity of the training path against plain fp32/fp64 torch restatements of the same ops (the reference
arithmetic of models/recnet.py ConvLayer / Conv4Channel / rectification lines and models/trainer.py losses), on the same
seeded inputs. Tolerances are stated per assert: fp16 hi+lo forward operands (~2^-21) with fp16 weights, fp32
accumulation; bf16 only where the device stores bf16 (dz, weight-gradient operands, loss Gram differences)."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from ffr_net_b200 import _lib
from h9 import fold_h9, from_h9, hilo_cat, own_rows, own_to_h9, rel_l2, to_h9

pytestmark = pytest.mark.gpu
EPI = _lib.EPI
P = _lib.ptr
TAPS9 = (ctypes.c_int * 9)(*[(r - 1) * 9 + (s - 1) for r in range(3) for s in range(3)])


@pytest.fixture(params=["rowmajor", "pixmajor"])
def tile_mode(request, lib):
    lib.ffr_debug_set_pixmajor(1 if request.param == "pixmajor" else 0)
    yield request.param
    lib.ffr_debug_set_pixmajor(-1)


def _scatter(dev="cuda", off=0):
    from ffr_net_b200.recnet import _h9_scatter
    return _h9_scatter(off, dev)


def _conv_fwd(lib, a_h, lo_off, w16, cin_p, cout_p, n, z, part, pix):
    d = _lib.ConvGemmDesc()
    d.a, d.a_rows, d.a_cols, d.a_ld = P(a_h), n * 81, a_h.shape[1], a_h.shape[1]
    d.wp, d.Cin, d.Cout, d.ntaps = P(w16), cin_p, cout_p, 9
    d.tap_row_shift = ctypes.cast(TAPS9, ctypes.c_void_p)
    d.M, d.rows_per_img, d.Wp, d.S, d.h0, d.n_img = n * 81, 81, 9, 7, 1, n
    d.flags = EPI.GEOM | EPI.STATS | EPI.OUT_F32 | (EPI.PIXMAJOR if pix else 0)
    d.out_f32, d.stats_part, d.num_splits = P(z), P(part), 1
    d.a_hilo, d.a_lo_off, d.f16 = (1 if lo_off else 0), lo_off, 1
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), _lib.stream_ptr()), "conv fwd")


@pytest.mark.parametrize("n,cin,cout,groups", [(3, 128, 64, 1), (5, 561, 256, 1), (64, 64, 49, 2), (130, 192, 128, 1)])
def test_conv_fwd_f16_hilo_stats_finalize(lib, tile_mode, n, cin, cout, groups):
    """fp16 hi+lo activations x fp16 weights -> fp32 z, deterministic BatchNorm partial sums, ffr_bn_finalize (per-group
    mean / rstd, running statistics in call order) vs float64 torch."""
    g = torch.Generator().manual_seed(n + cin)
    cin_p, cout_p = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    x = torch.randn(n, cin, 7, 7, generator=g) * 1.7
    w = torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)
    pix = tile_mode == "pixmajor"
    a_h = hilo_cat(to_h9(x, cin_p)).cuda()
    w16 = torch.empty(cout_p, 9 * cin_p, dtype=torch.float16, device="cuda")
    wt = torch.empty(cin_p, 9 * cout_p, dtype=torch.bfloat16, device="cuda")
    wd = w.cuda()
    _lib.check(lib.ffr_pack_conv3x3_f16(P(wd), cout, cin, cout_p, cin_p, P(w16), P(wt), _lib.stream_ptr()))
    m_tiles = 49 * ((n + 127) // 128) if pix else (n * 81 + 127) // 128
    part = torch.full((4 * m_tiles, 2, cout_p), 7.0, device="cuda")
    z = torch.full((n * 81, cout_p), 7.0, device="cuda")
    _conv_fwd(lib, a_h, cin_p, w16, cin_p, cout_p, n, z, part, pix)
    torch.cuda.synchronize()
    z_ref = F.conv2d(F.pad(x.double(), (1, 1, 1, 1), mode="reflect"), w.half().double())
    got = from_h9(z.cpu(), cout_p)
    assert rel_l2(got[:, :cout], z_ref) <= 2e-5                      # x = hi + lo to 2^-21, weights exact in fp16
    assert got[:, cout:].abs().max().item() == 0.0 if cout_p > cout else True
    s = part.cpu().double().sum(0)                                  # [2][cout_p]
    # tcgen05 adds the products of one instruction into the fp32 accumulator with truncation: a systematic ~1e-5 shrink
    assert rel_l2(s[0, :cout], z_ref.sum((0, 2, 3))) <= 2e-4
    assert rel_l2(s[1, :cout], (z_ref ** 2).sum((0, 2, 3))) <= 1e-4
    # finalize
    npg = n // groups
    rm0, rv0 = torch.rand(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    rm, rv = rm0.clone().cuda(), rv0.clone().cuda()
    nbt = torch.tensor(3, dtype=torch.int64, device="cuda")
    mr = torch.empty(groups, 2, cout_p, device="cuda")
    _lib.check(lib.ffr_bn_finalize(P(part), part.shape[0], 1 if pix else 0, n, npg, cout_p, cout, 0.1, 1e-5, P(rm), P(rv),
                                   P(nbt), P(mr), _lib.stream_ptr()))
    torch.cuda.synchronize()
    rm_ref, rv_ref = rm0.double(), rv0.double()
    for gi in range(groups):
        zg = z_ref[gi * npg:(gi + 1) * npg]
        mean, var = zg.mean((0, 2, 3)), zg.var((0, 2, 3), unbiased=False)
        cnt = npg * 49
        assert rel_l2(mr[gi, 0, :cout].cpu(), mean) <= 1e-4
        assert rel_l2(mr[gi, 1, :cout].cpu(), 1 / torch.sqrt(var + 1e-5)) <= 5e-5
        rm_ref = 0.9 * rm_ref + 0.1 * mean
        rv_ref = 0.9 * rv_ref + 0.1 * var * cnt / (cnt - 1)
    assert rel_l2(rm.cpu(), rm_ref) <= 1e-5 and rel_l2(rv.cpu(), rv_ref) <= 1e-5
    assert int(nbt) == 3 + groups


@pytest.mark.parametrize("n,c,groups,with_res", [(4, 64, 1, True), (64, 128, 2, False), (64, 49, 2, True)])
def test_bn_act_fwd_bwd(lib, n, c, groups, with_res):
    """ffr_bn_act_fwd / ffr_bn_act_bwd (batch-statistics BatchNorm + PReLU + residual, per group) vs fp64 autograd."""
    g = torch.Generator().manual_seed(n * 7 + c)
    cp = (c + 63) // 64 * 64
    npg = n // groups
    z = torch.randn(n, c, 7, 7, generator=g) * 2 + 0.3
    res = torch.randn(n, c, 7, 7, generator=g)
    gam = torch.empty(c).uniform_(0.5, 1.5, generator=g)
    bet = torch.empty(c).uniform_(-0.3, 0.3, generator=g)
    slo = torch.empty(c).uniform_(0.1, 0.4, generator=g)
    da_grid = torch.randn(n * 81, cp, generator=g) * 0.1            # gradient on the padded grid (own + mirror rows)
    dadd = torch.randn(n, c, 7, 7, generator=g) * 0.1
    dv = torch.randn(n, 512, generator=g)
    # ---- reference (fp64) ----
    zr = z.double().requires_grad_(True)
    gr, br, sr = gam.double().requires_grad_(True), bet.double().requires_grad_(True), slo.double().requires_grad_(True)
    outs = []
    for gi in range(groups):
        zg = zr[gi * npg:(gi + 1) * npg]
        y = F.batch_norm(zg, None, None, gr, br, True, 0.1, 1e-5)
        outs.append(F.prelu(y, sr))
    a_ref = torch.cat(outs) + (res.double() if with_res else 0)
    A = fold_h9(da_grid, c, n).double() + dadd.double() + dv[:, :c].double().view(n, c, 1, 1) / 49
    a_ref.backward(A)
    # ---- device ----
    zd = own_to_h9(z, cp).cuda()
    mr = torch.zeros(groups, 2, cp)
    for gi in range(groups):
        zg = z[gi * npg:(gi + 1) * npg].double()
        mr[gi, 0, :c] = zg.mean((0, 2, 3)).float()
        mr[gi, 1, :c] = (1 / torch.sqrt(zg.var((0, 2, 3), unbiased=False) + 1e-5)).float()
        mr[gi, 1, c:] = 316.0
    mr = mr.cuda()
    res_h = hilo_cat(to_h9(res, cp)).cuda()
    out_h = torch.zeros(n * 81, 2 * cp, dtype=torch.float16, device="cuda")
    out_b = torch.zeros(n * 81, cp, dtype=torch.bfloat16, device="cuda")
    out_f = torch.zeros(n * 81, cp, device="cuda")
    tab = _scatter()
    st = _lib.stream_ptr()
    gd, bd, sd = gam.cuda(), bet.cuda(), slo.cuda()
    _lib.check(lib.ffr_bn_act_fwd(P(zd), cp, P(mr), P(gd), P(bd), P(sd), P(res_h) if with_res else None,
                                  2 * cp if with_res else 0, cp if with_res else 0, P(out_h), 2 * cp, cp, P(out_b), cp,
                                  P(out_f), cp, 0, P(tab), 4, n, npg, cp, c, st))
    torch.cuda.synchronize()
    a32 = a_ref.detach().float()
    exp_rows = to_h9(a32, cp)                                       # own rows + reflection mirrors
    got = out_h[:, :cp].float() + out_h[:, cp:].float()
    assert (got.cpu() - exp_rows).abs().max().item() <= 2e-5 * max(1.0, exp_rows.abs().max().item())
    assert rel_l2(out_b.float().cpu(), exp_rows) <= 4e-3
    assert rel_l2(from_h9(out_f.cpu(), c), a32) <= 1e-5
    # backward
    G = groups
    rows = lib.ffr_bn_act_bwd_partial_rows(G, cp)
    partial = torch.zeros(rows * 3 * cp, device="cuda")
    gsum = torch.zeros(G * 2 * cp, device="cuda")
    afold = torch.zeros(n * 81, cp, device="cuda")
    dz = torch.full((n * 81, cp), 3.0, dtype=torch.bfloat16, device="cuda")
    dg, db, dsl = (torch.full((c,), 9.0, device="cuda") for _ in range(3))
    dadd_h9 = own_to_h9(dadd, cp).cuda()
    da_d, dv_d = da_grid.cuda(), dv.cuda()                          # (device tensors must outlive the launch)
    _lib.check(lib.ffr_bn_act_bwd(P(da_d), cp, 0, P(tab), 4, P(dadd_h9), cp, 0, P(dv_d), 512, 1.0 / 49,
                                  P(zd), cp, P(mr), P(gd), P(bd), P(sd), P(afold), cp, P(partial), P(gsum), P(dg), P(db),
                                  P(dsl), 0, c, P(dz), cp, n, npg, cp, st))
    torch.cuda.synchronize()
    assert rel_l2(from_h9(afold.cpu(), c), A) <= 1e-5
    assert rel_l2(dg.cpu(), gr.grad) <= 2e-5 and rel_l2(db.cpu(), br.grad) <= 2e-5 and rel_l2(dsl.cpu(), sr.grad) <= 2e-5
    assert rel_l2(from_h9(dz.float().cpu(), c), zr.grad) <= 4e-3    # one bf16 rounding of the final value
    halo = torch.ones(81, dtype=torch.bool)
    halo[own_rows()] = False
    assert dz.view(n, 81, cp)[:, halo.cuda()].float().abs().max().item() == 0.0


@pytest.mark.parametrize("n,cin,cout", [(3, 128, 64), (4, 561, 256), (130, 256, 128)])
def test_wgrad_and_dgrad(lib, tile_mode, n, cin, cout):
    """ffr_wgrad (bf16 operands, deterministic slabs and accumulate) and the dgrad GEMM with fp32 output vs autograd of the
    reflection-padded convolution."""
    g = torch.Generator().manual_seed(cin)
    cin_p, cout_p = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
    x = torch.randn(n, cin, 7, 7, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).bfloat16().float()
    dzv = torch.randn(n, cout, 7, 7, generator=g).bfloat16().float()
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(F.pad(xr, (1, 1, 1, 1), mode="reflect"), wr).backward(dzv.double())
    x_b = to_h9(x, cin_p).bfloat16().cuda()
    dz_b = to_h9(dzv, cout_p, mirror=False).bfloat16().cuda()
    st = _lib.stream_ptr()
    for det in (0, 1):
        ws = torch.zeros(int(lib.ffr_wgrad_workspace_floats(n * 81, cout, cin, 9, det)), device="cuda")
        dw = torch.full((cout, cin, 3, 3), 7.0, device="cuda")
        _lib.check(lib.ffr_wgrad(P(dz_b), cout_p, P(x_b), cin_p, 0, n * 81, cout, cin, 9, 0, det, 0, cin, -1, P(dw), None,
                                 P(ws), st))
        torch.cuda.synchronize()
        assert rel_l2(dw.cpu(), wr.grad) <= 1e-5, det
        if det:
            dw2 = dw.clone()
            _lib.check(lib.ffr_wgrad(P(dz_b), cout_p, P(x_b), cin_p, 0, n * 81, cout, cin, 9, 0, 1, 1, cin, -1, P(dw2), None,
                                     P(ws), st))                     # accumulate: exactly twice the gradient
            torch.cuda.synchronize()
            assert torch.equal(dw2, dw * 2)
    w16 = torch.empty(cout_p, 9 * cin_p, dtype=torch.float16, device="cuda")
    wt = torch.empty(cin_p, 9 * cout_p, dtype=torch.bfloat16, device="cuda")
    wd = w.cuda()
    _lib.check(lib.ffr_pack_conv3x3_f16(P(wd), cout, cin, cout_p, cin_p, P(w16), P(wt), st))
    pix = tile_mode == "pixmajor"
    dx = torch.zeros(n * 81, cin_p, device="cuda")
    d = _lib.ConvGemmDesc()
    d.a, d.a_rows, d.a_cols, d.a_ld = P(dz_b), n * 81, cout_p, cout_p
    d.wp, d.Cin, d.Cout, d.ntaps = P(wt), cout_p, cin_p, 9
    d.tap_row_shift = ctypes.cast(TAPS9, ctypes.c_void_p)
    d.M, d.rows_per_img, d.Wp, d.S, d.h0, d.n_img = n * 81, 81, 9, 7, 1, n
    d.flags = EPI.OUT_F32 | ((EPI.PIXMAJOR | EPI.PIX_DGRAD) if pix else 0)
    d.out_f32, d.num_splits = P(dx), 1
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dgrad")
    torch.cuda.synchronize()
    assert rel_l2(fold_h9(dx.cpu(), cin, n), xr.grad) <= 1e-5


def test_wgrad_one_tap_with_bias_column(lib):
    """ntaps = 1: dW[j][k] = sum_rows Y[row][j] X[row][k], ones column -> bias gradient (Linear weight gradients)."""
    g = torch.Generator().manual_seed(5)
    rows = 3 * 512
    y = torch.randn(rows, 512, generator=g).bfloat16()
    x = torch.zeros(rows, 64).bfloat16()
    x[:, :32] = torch.randn(rows, 32, generator=g).bfloat16()
    x[:, 32] = 1.0
    ws = torch.zeros(int(lib.ffr_wgrad_workspace_floats(rows, 512, 33, 1, 1)), device="cuda")
    dw = torch.full((512, 32), 5.0, device="cuda")
    db = torch.full((512,), 5.0, device="cuda")
    yd, xd = y.cuda(), x.cuda()
    _lib.check(lib.ffr_wgrad(P(yd), 512, P(xd), 64, 0, rows, 512, 33, 1, 0, 1, 0, 32, 32, P(dw), P(db), P(ws),
                             _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = y.double().t() @ x.double()
    assert rel_l2(dw.cpu(), ref[:, :32]) <= 1e-5 and rel_l2(db.cpu(), ref[:, 32]) <= 1e-5


# ----------------------------------------------------------------------------------------------------------
def _chan_params(g):
    p = {"w0": torch.randn(32, 561, generator=g) * (2 / 561) ** 0.5, "b0": torch.empty(32).uniform_(-0.1, 0.1, generator=g),
         "w8": torch.randn(512, 32, generator=g) * (2 / 32) ** 0.5, "b8": torch.empty(512).uniform_(-0.1, 0.1, generator=g)}
    for a, b in ((2, 3), (5, 6)):
        p["w%d" % a] = torch.randn(512, 32, generator=g) * (2 / 32) ** 0.5
        p["b%d" % a] = torch.empty(512).uniform_(-0.1, 0.1, generator=g)
        p["w%d" % b] = torch.randn(32, 512, generator=g) * (2 / 512) ** 0.5
        p["b%d" % b] = torch.empty(32).uniform_(-0.1, 0.1, generator=g)
    for i in (1, 4, 7):
        p["s%d" % i] = torch.empty(512).uniform_(0.1, 0.4, generator=g)
    return p


def _chan_ref(x, p):
    """Conv4Channel (recnet.py:372-386) + rectification (:406,:410) in the given dtype; returns intermediates."""
    n = x.shape[0]
    flat = x.reshape(n, 512, 49)
    xh = F.normalize(flat, dim=2)
    ss_c = torch.bmm(xh, xh.transpose(1, 2))
    h = F.linear(torch.cat((flat, ss_c), 2), p["w0"], p["b0"])
    g0 = h
    h = F.prelu(h, p["s1"])
    h = F.linear(F.linear(h, p["w2"], p["b2"]), p["w3"], p["b3"])
    g1 = h
    h = F.prelu(h, p["s4"])
    h = F.linear(F.linear(h, p["w5"], p["b5"]), p["w6"], p["b6"])
    g2 = h
    h7 = F.prelu(h, p["s7"])
    m = torch.sigmoid(F.linear(h7, p["w8"], p["b8"]))
    fc = torch.matmul(m, flat)
    return dict(g0=g0, g1=g1, g2=g2, h7=h7, m=m, fc=fc)


def _run_prep(lib, x, p, hilo=True):
    n = x.shape[0]
    dev = "cuda"
    pd = {k: v.cuda().contiguous() for k, v in p.items()}
    A = torch.zeros(2112, device=dev)
    w8t = torch.zeros(64, 512, dtype=torch.bfloat16, device=dev)
    st = _lib.stream_ptr()
    _lib.check(lib.ffr_chan_compose(P(pd["w2"]), P(pd["b2"]), P(pd["w3"]), P(pd["b3"]), P(pd["w5"]), P(pd["b5"]), P(pd["w6"]),
                                    P(pd["b6"]), P(A[0:]), P(A[1024:]), P(A[1056:]), P(A[2080:]), P(pd["w8"]), P(w8t), st))
    R = n * 81
    mk = lambda c: (torch.zeros(R, 2 * c, dtype=torch.float16, device=dev), torch.zeros(R, c, dtype=torch.bfloat16, device=dev))
    s0, cm, fm = mk(576), mk(1536), mk(1024)
    o = dict(g=[torch.zeros(n * 512, 32, device=dev) for _ in range(3)],
             h7b=torch.zeros(n * 512, 64, dtype=torch.bfloat16, device=dev),
             xk=torch.zeros(n * 512, 64, dtype=torch.bfloat16, device=dev),
             mch2=torch.zeros(n * 512, 1024, dtype=torch.float16, device=dev),
             x3=torch.zeros(n * 64, 1536, dtype=torch.float16, device=dev),
             fcraw=torch.zeros(n * 512, 64, device=dev),
             inv_c=torch.zeros(n * 512, device=dev), tmat=torch.zeros(n, 49, 32, device=dev),
             ss=torch.zeros(n, 49, 49, device=dev), s0=s0, cm=cm, fm=fm, A=A, w8t=w8t, pd=pd)
    d = _lib.PrepTrainDesc()
    xd = x.cuda().contiguous()
    o["x"] = xd
    d.x, d.w0, d.b0 = P(xd), P(pd["w0"]), P(pd["b0"])
    d.slope1, d.slope4, d.slope7 = P(pd["s1"]), P(pd["s4"]), P(pd["s7"])
    d.A1, d.c1, d.A2, d.c2 = P(A[0:]), P(A[1024:]), P(A[1056:]), P(A[2080:])
    d.w8, d.b8 = P(pd["w8"]), P(pd["b8"])
    for nm, (h, b), c in (("s0", s0, 576), ("cm", cm, 1536)):
        setattr(d, nm + "_h", P(h)); setattr(d, nm + "_ld", 2 * c); setattr(d, nm + "_lo", c)
        setattr(d, nm + "_b", P(b)); setattr(d, nm + "_ldb", c)
    d.g0, d.g1, d.g2 = P(o["g"][0]), P(o["g"][1]), P(o["g"][2])
    d.h7b, d.xk, d.mch2, d.x3 = P(o["h7b"]), P(o["xk"]), P(o["mch2"]), P(o["x3"])
    d.inv_c, d.tmat, d.ss_space = P(o["inv_c"]), P(o["tmat"]), P(o["ss"])
    _lib.check(lib.ffr_recnet_prep_train(ctypes.byref(d), n, st), "prep_train")
    g = _lib.ConvGemmDesc()                # feat_channel = M_channel @ X on the tcgen05 GEMM (K = 1536, fp16 hi/lo splits)
    g.a, g.a_rows, g.a_cols, g.a_ld = P(o["mch2"]), n * 512, 1024, 1024
    g.wp, g.Cin, g.Cout, g.ntaps = P(o["x3"]), 512, 64, 3
    choff = (ctypes.c_int * 3)(0, 0, 512)
    g.tap_ch_off = ctypes.cast(choff, ctypes.c_void_p)
    g.M, g.flags, g.out_f32 = n * 512, EPI.OUT_F32, P(o["fcraw"])
    g.num_splits, g.b_rows_per_mtile, g.b_mtile_div, g.f16 = 1, 64, 4, 1
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(g), st), "feat_channel GEMM")
    _lib.check(lib.ffr_fc_scatter(P(o["fcraw"]), P(fm[0]), 2048, 1024, P(fm[1]), 1024, n, st), "fc_scatter")
    torch.cuda.synchronize()
    return o


def test_recnet_prep_train(lib):
    """selfSimilarity, cat fan-outs, Conv4Channel and feat_channel (+ flip / cat) of the training forward vs fp64 torch."""
    g = torch.Generator().manual_seed(11)
    n = 3
    x = torch.randn(n, 512, 7, 7, generator=g) * 0.8
    p = _chan_params(g)
    o = _run_prep(lib, x, p)
    ref = _chan_ref(x.double(), {k: v.double() for k, v in p.items()})
    hl = lambda t, c: (t[0][:, :c].float() + t[0][:, c:].float()).cpu()
    # composed maps
    A = o["A"].cpu().double()
    assert rel_l2(A[:1024].view(32, 32), p["w3"].double() @ p["w2"].double()) <= 1e-5
    assert rel_l2(A[1024:1056], p["w3"].double() @ p["b2"].double() + p["b3"].double()) <= 1e-5
    assert rel_l2(o["w8t"][:32].float().cpu(), p["w8"].t().bfloat16().float()) == 0.0
    # chain
    for i, k in enumerate(("g0", "g1", "g2")):
        assert rel_l2(o["g"][i].view(n, 512, 32).cpu(), ref[k]) <= 2e-5, k
    assert rel_l2(o["h7b"][:, :32].float().view(n, 512, 32).cpu(), ref["h7"]) <= 4e-3
    assert torch.equal(o["h7b"][:, 32].float().cpu(), torch.ones(n * 512)) and o["h7b"][:, 33:].float().abs().max().item() == 0
    m_dev = (o["mch2"][:, :512].float() + o["mch2"][:, 512:].float()).view(n, 512, 512).cpu()
    assert rel_l2(m_dev, ref["m"]) <= 2e-5
    x3 = o["x3"].float().view(n, 64, 3, 512).cpu()
    assert rel_l2((x3[:, :49, 0] + x3[:, :49, 1]).transpose(1, 2), x.reshape(n, 512, 49)) <= 1e-6
    assert torch.equal(x3[:, :, 0], x3[:, :, 2]) and x3[:, 49:].abs().max().item() == 0
    assert rel_l2(o["fcraw"][:, :49].view(n, 512, 49).cpu(), ref["fc"]) <= 2e-5
    flat = x.reshape(n, 512, 49)
    assert rel_l2(o["xk"][:, :49].float().view(n, 512, 49).cpu(), flat) <= 4e-3 and o["xk"][:, 49:].float().abs().max().item() == 0
    assert rel_l2(o["inv_c"].view(n, 512).cpu(), 1 / flat.double().norm(dim=2)) <= 1e-5
    # Conv4Space input: X | ss_space | 0 with reflection halo
    xs = F.normalize(flat.double().transpose(1, 2), dim=2)
    ss_space = torch.bmm(xs, xs.transpose(1, 2)).reshape(n, 49, 7, 7)
    assert rel_l2(o["ss"].cpu().view(n, 49, 7, 7), ss_space) <= 1e-5
    s0_ref = to_h9(torch.cat((x.double(), ss_space), 1).float(), 576)
    assert (hl(o["s0"], 576) - s0_ref).abs().max().item() <= 1e-5
    assert rel_l2(o["s0"][1].float().cpu(), s0_ref) <= 4e-3
    assert (hl(o["cm"], 1536)[:, 1024:] - to_h9(x, 512)).abs().max().item() <= 1e-5
    fc = ref["fc"].reshape(n, 512, 7, 7).float()
    fm_ref = to_h9(torch.cat((torch.flip(fc, [3]), fc), 1), 1024)
    assert rel_l2(hl(o["fm"], 1024), fm_ref) <= 2e-5
    assert rel_l2(o["fm"][1].float().cpu(), fm_ref) <= 4e-3


def test_channel_rectifier_backward(lib):
    """fc_bwd_gather -> dM_pre GEMM -> dh7 GEMM -> dW8 wgrad -> chan_bwd -> compose_bwd vs fp64 autograd of the Conv4Channel
    chain (the gradient enters on the padded grid of the ChannelFlipMerge input)."""
    g = torch.Generator().manual_seed(12)
    n = 2
    x = torch.randn(n, 512, 7, 7, generator=g) * 0.8
    p = _chan_params(g)
    dfm = torch.randn(n * 81, 1024, generator=g) * 0.05
    pr = {k: v.double().requires_grad_(True) for k, v in p.items()}
    ref = _chan_ref(x.double(), pr)
    fc = ref["fc"].reshape(n, 512, 7, 7)
    fm = torch.cat((torch.flip(fc, [3]), fc), 1)
    pad = F.pad(fm, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1).reshape(n * 81, 1024)
    (pad * dfm.double()).sum().backward()
    o = _run_prep(lib, x, p)
    pd = o["pd"]
    st = _lib.stream_ptr()
    dev = "cuda"
    dfc_op = torch.zeros(n * 512, 64, dtype=torch.bfloat16, device=dev)
    dfm_d = dfm.cuda()
    _lib.check(lib.ffr_fc_bwd_gather(P(dfm_d), 1024, P(dfc_op), n, st))
    dmpre = torch.zeros(n * 512, 512, dtype=torch.bfloat16, device=dev)
    d = _lib.ConvGemmDesc()
    d.a, d.a_rows, d.a_cols, d.a_ld = P(dfc_op), n * 512, 64, 64
    d.wp, d.Cin, d.Cout, d.ntaps = P(o["xk"]), 64, 512, 1
    d.M, d.flags = n * 512, EPI.MUL_DSIG
    d.flags |= EPI.RES_F16
    d.out, d.ldo, d.res, d.ldres = P(dmpre), 512, P(o["mch2"]), 1024
    d.num_splits, d.b_rows_per_mtile, d.b_mtile_div = 1, 512, 4
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dM_pre")
    dh7 = torch.zeros(n * 512, 64, device=dev)
    d = _lib.ConvGemmDesc()
    d.a, d.a_rows, d.a_cols, d.a_ld = P(dmpre), n * 512, 512, 512
    d.wp, d.Cin, d.Cout, d.ntaps = P(o["w8t"]), 512, 64, 1
    d.M, d.flags, d.out_f32, d.num_splits = n * 512, EPI.OUT_F32, P(dh7), 1
    _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dh7")
    gr = {k: torch.full_like(v, 3.0).cuda() for k, v in p.items()}
    ws = torch.zeros(int(lib.ffr_wgrad_workspace_floats(n * 512, 512, 33, 1, 1)), device=dev)
    _lib.check(lib.ffr_wgrad(P(dmpre), 512, P(o["h7b"]), 64, 0, n * 512, 512, 33, 1, 0, 1, 0, 32, 32, P(gr["w8"]), P(gr["b8"]),
                             P(ws), st))
    part = torch.zeros(n * lib.ffr_chan_bwd_part_floats(), device=dev)
    dsp = torch.zeros(n * 3 * 512, device=dev)
    tmp = torch.zeros(2112, device=dev)
    A = o["A"]
    _lib.check(lib.ffr_chan_bwd(P(o["x"]), P(dh7), P(o["g"][0]), P(o["g"][1]), P(o["g"][2]), P(o["inv_c"]), P(o["tmat"]),
                                P(A[0:]), P(A[1056:]), P(pd["s1"]), P(pd["s4"]), P(pd["s7"]), P(part), P(dsp), P(tmp),
                                P(gr["b0"]), P(gr["w0"]), P(gr["s1"]), P(gr["s4"]), P(gr["s7"]), 0, n, st))
    _lib.check(lib.ffr_chan_compose_bwd(P(pd["w2"]), P(pd["b2"]), P(pd["w3"]), P(pd["w5"]), P(pd["b5"]), P(pd["w6"]),
                                        P(tmp[1056:]), P(tmp[2080:]), P(tmp[0:]), P(tmp[1024:]),
                                        P(gr["w2"]), P(gr["b2"]), P(gr["w3"]), P(gr["b3"]), P(gr["w5"]), P(gr["b5"]),
                                        P(gr["w6"]), P(gr["b6"]), 0, st))
    torch.cuda.synchronize()
    errs = {k: rel_l2(gr[k].cpu(), pr[k].grad) for k in p}
    print("channel rectifier backward, rel L2 per parameter:", {k: "%.2e" % v for k, v in errs.items()})
    # bf16 operands (dFC, X rows, M_channel, dM_pre, h7) on the way: 1e-2 per tensor
    assert max(errs.values()) <= 1e-2, errs


def test_feat_space_fwd_bwd(lib):
    g = torch.Generator().manual_seed(13)
    n = 3
    x = torch.randn(n, 512, 7, 7, generator=g)
    pre = torch.randn(n, 49, 49, generator=g)                        # [n][i][j]
    dcm = torch.randn(n * 81, 1024, generator=g) * 0.1
    dfs = torch.randn(n, 512, 7, 7, generator=g) * 0.1
    prr = pre.double().requires_grad_(True)
    M = torch.sigmoid(prr)
    fs = torch.matmul(x.double().reshape(n, 512, 49), M).reshape(n, 512, 7, 7)
    pad = F.pad(fs, (1, 1, 1, 1), mode="reflect").permute(0, 2, 3, 1).reshape(n * 81, 512)
    ((pad * dcm[:, :512].double()).sum() + (fs * dfs.double()).sum()).backward()
    msp = torch.zeros(n * 81, 64)
    msp.view(n, 81, 64)[:, own_rows(), :49] = M.detach().float().permute(0, 2, 1)        # row = pixel j, column = i
    msp = msp.cuda()
    cm_h = torch.zeros(n * 81, 2 * 1536, dtype=torch.float16, device="cuda")
    cm_b = torch.zeros(n * 81, 1536, dtype=torch.bfloat16, device="cuda")
    fs_f = torch.zeros(n * 81, 512, device="cuda")
    st = _lib.stream_ptr()
    xd = x.cuda()
    _lib.check(lib.ffr_feat_space_train(P(xd), P(msp), P(cm_h), 3072, 1536, P(cm_b), 1536, P(fs_f), 512, n, st))
    dfs_h9, dcm_d = own_to_h9(dfs).cuda(), dcm.cuda()
    dmsp = torch.full((n * 81, 64), 4.0, device="cuda")
    _lib.check(lib.ffr_feat_space_bwd(P(xd), P(msp), P(dcm_d), 1024, P(dfs_h9), 512, P(dmsp), n, st))
    torch.cuda.synchronize()
    fs32 = fs.detach().float()
    assert rel_l2(from_h9(fs_f.cpu(), 512), fs32) <= 1e-5
    got = (cm_h[:, :512].float() + cm_h[:, 1536:2048].float()).cpu()
    assert rel_l2(got, to_h9(fs32, 512)) <= 1e-5
    assert rel_l2(cm_b[:, :512].float().cpu(), to_h9(fs32, 512)) <= 4e-3
    dpre = dmsp.cpu().view(n, 81, 64)[:, own_rows(), :49].permute(0, 2, 1)              # -> [n][i][j]
    assert rel_l2(dpre, prr.grad) <= 1e-4
    assert dmsp.cpu().view(n, 81, 64)[:, own_rows(), 49:].abs().max().item() == 0.0


# ----------------------------------------------------------------------------------------------------------
def _selfsim_ref(xt, f):
    """(ss_space, ss_channel) MSE pieces of trainer.py:157-165 for one call: sums of squared Gram differences."""
    def grams(t):
        n = t.shape[0]
        v = t.reshape(n, 512, 49)
        a = F.normalize(v.transpose(1, 2), dim=2)
        b = F.normalize(v, dim=2)
        return torch.bmm(a, a.transpose(1, 2)), torch.bmm(b, b.transpose(1, 2))
    ts, tc = grams(xt)
    gs, gc = grams(f)
    return ((ts - gs) ** 2), ((tc - gc) ** 2)


def test_selfsim_losses(lib):
    """Fused self-similarity losses (channel: tcgen05 Gram-difference GEMM + gradient GEMM; space: SIMT) vs fp64 autograd,
    two groups of 2 samples sharing the targets of group 0."""
    from ffr_net_b200 import losses
    g = torch.Generator().manual_seed(21)
    n = 2
    xt = torch.randn(n, 512, 7, 7, generator=g)
    fc = torch.randn(2 * n, 512, 7, 7, generator=g) * 1.3 + 0.2
    fs = torch.randn(2 * n, 512, 7, 7, generator=g) * 0.7
    w0 = 0.8
    fcr, fsr = fc.double().requires_grad_(True), fs.double().requires_grad_(True)
    tgt = torch.cat((xt, xt)).double()
    _, dc = _selfsim_ref(tgt, fcr)
    ds, _ = _selfsim_ref(tgt, fsr)
    # L1 = 1/2 (1/2 (ms_non + ms_ocl) + 1/2 (mc_non + mc_ocl)), each MSE a mean over n samples
    l_c = sum(dc[i * n:(i + 1) * n].mean() for i in range(2)) / 2
    l_s = sum(ds[i * n:(i + 1) * n].mean() for i in range(2)) / 2
    (w0 * 0.5 * (l_s + l_c)).backward()
    lw = losses.LossWorkspace(n, "cuda")
    fc_h9, fs_h9 = own_to_h9(fc).cuda(), own_to_h9(fs).cuda()
    xd = xt.cuda()
    losses.selfsim_channel(lw, fc_h9, xd, 2 * n, n, 0, w0)
    losses.selfsim_space(lw, fs_h9, xd, 2 * n, n, 0, w0)
    torch.cuda.synchronize()
    chan = lw.chan_sums.view(2, 64).sum(1).cpu().double()
    assert rel_l2(chan, torch.stack([dc[i * n:(i + 1) * n].sum() for i in range(2)]).detach()) <= 1e-3
    assert rel_l2(lw.space_part.cpu(), ds.sum((1, 2)).detach()) <= 1e-5
    assert rel_l2(from_h9(lw.dfs.cpu(), 512), fsr.grad) <= 1e-4
    assert rel_l2(from_h9(lw.dfc.cpu(), 512), fcr.grad) <= 1e-2      # D and F^ enter the gradient GEMM as single bf16


def test_triplet_identity_and_finalize(lib):
    from ffr_net_b200 import losses
    g = torch.Generator().manual_seed(22)
    n = 37
    f_non, f_ocl = torch.randn(n, 512, generator=g), torch.randn(n, 512, generator=g)
    e_non, e_ocl = F.normalize(torch.randn(n, 512, generator=g)), F.normalize(torch.randn(n, 512, generator=g))
    e_ocl[::3] = F.normalize(f_ocl[::3] + 0.3 * torch.randn(n, 512, generator=g)[::3])      # some inactive triplets
    w = [0.7, 1.3, 0.9, 1.1]
    fn, fo = f_non.double().requires_grad_(True), f_ocl.double().requires_grad_(True)
    pos = 1 - (F.normalize(fo) * F.normalize(e_non.double())).sum(1)
    neg = 1 - (F.normalize(fo) * F.normalize(e_ocl.double())).sum(1)
    trip = F.relu(pos - neg + 0.1).mean()
    ident = (F.mse_loss(fn, e_non.double()) + F.mse_loss(fo, e_non.double())) / 2
    (w[1] * trip + w[2] * ident).backward()
    lw = losses.LossWorkspace(n, "cuda")
    dev_in = [t.cuda() for t in (f_non, f_ocl, e_non, e_ocl)]
    losses.triplet_identity(lw, *dev_in, w[1], w[2])
    lw.space_part.copy_(torch.rand(2 * n, generator=g))
    lw.chan_sums.copy_(torch.rand(128, generator=g))
    ce = torch.tensor([2.5, 3.5], device="cuda")
    losses.finalize(lw, ce, w)
    torch.cuda.synchronize()
    assert rel_l2(lw.dvl[:n].cpu(), fn.grad) <= 1e-5 and rel_l2(lw.dvl[n:].cpu(), fo.grad) <= 1e-5
    out = lw.out.cpu().double()
    l_s = lw.space_part.cpu().double().sum() / (n * 2401) / 2
    l_c = lw.chan_sums.cpu().double().sum() / (n * 512 * 512) / 2
    exp = [w[0] * 0.5 * (l_s + l_c), w[1] * trip.item(), w[2] * ident.item(), w[3] * (2.5 / (1e-8 + w[3]) + 3.5),
           pos.mean().item(), neg.mean().item()]
    for i, e in enumerate(exp):
        assert abs(out[i].item() - float(e)) <= 1e-5 * max(1.0, abs(float(e))), (i, out[i].item(), float(e))
    assert abs(out[6].item() - float(sum(exp[:4]))) <= 1e-4


def test_grouped_head_matches_oracle(lib):
    """GroupedHead: one cosine GEMM for both calls, per-call mean CE, gradients vs fp64 autograd of AddMarginProduct + CE."""
    from oracle import recnet as orr
    from ffr_net_b200 import head
    from ffr_net_b200.recnet import AddMarginProduct
    g = torch.Generator().manual_seed(23)
    n, classes = 32, 10575
    cls = AddMarginProduct(512, classes).cuda()
    v = torch.randn(2 * n, 512, generator=g)
    label = torch.randint(0, classes, (n,), generator=g)
    gl = torch.tensor([0.9, 1.4])
    wr, vr = cls.weight.detach().cpu().double().requires_grad_(True), v.double().requires_grad_(True)
    l0 = F.cross_entropy(orr.add_margin_product({"classifier.weight": wr}, vr[:n], label)[0], label)
    l1 = F.cross_entropy(orr.add_margin_product({"classifier.weight": wr}, vr[n:], label)[0], label)
    (gl[0] * l0 + gl[1] * l1).backward()
    h = head.GroupedHead(cls, 2 * n, n, "cuda")
    ce = torch.zeros(2, device="cuda")
    vd = v.cuda()
    lab_d, gl_d = label.cuda(), gl.cuda()
    h.forward(vd, lab_d, ce)
    dw = torch.zeros(classes, 512, device="cuda")
    dv = h.backward(gl_d, dw)
    torch.cuda.synchronize()
    assert abs(ce[0].item() - l0.item()) <= 2e-4 * l0.item() and abs(ce[1].item() - l1.item()) <= 2e-4 * l1.item()
    assert rel_l2(dv.cpu(), vr.grad) <= 1e-2 and rel_l2(dw.cpu(), wr.grad) <= 1e-2
