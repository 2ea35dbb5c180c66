"""CPU tests of the host-side logic: the C-ABI library loads and exports every declared symbol, the drop-in modules
keep the reference's state_dict layout, and the weight folding / packing / layout logic (validated through a torch
emulation of the device algorithm) reproduces the oracle."""
import ctypes
import os

import pytest
import torch

import emulate
from oracle import backbone as ob
from oracle import recnet as orr
from ffr_net_b200 import _lib, layout, packing
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet, init_weights


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert len(names) >= 20 and "ffr_conv_gemm" in names and "ffr_threshold_sweep" in names
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    lib.ffr_version.restype = ctypes.c_int
    assert lib.ffr_version() >= 100
    assert set(names) <= set(_lib._SIGNATURES), set(names) - set(_lib._SIGNATURES)


def test_probe_library_is_separate_and_exports_its_header():
    """The hardware probes / micro-benchmarks (csrc/probe.cu) are NOT in the product library."""
    names = _lib.declared_symbols(_lib.PROBE_HEADER_PATH)
    assert set(names) == {"ffr_debug_rowshift_probe", "ffr_debug_mn_probe", "ffr_debug_mma_bench"}
    probe, main = ctypes.CDLL(_lib.PROBE_LIB_PATH), ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(probe, n) and not hasattr(main, n), n


def test_ctypes_signatures_match_the_headers():
    """Every ctypes signature in _lib._SIGNATURES has the arity and the argument kinds (pointer / int / int64 / float / u32)
    of its declaration in include/*.h — a float passed as an int would silently corrupt a call."""
    import re
    kinds = {ctypes.c_int: "i", ctypes.c_float: "f", ctypes.c_int64: "q", ctypes.c_uint32: "I", ctypes.c_longlong: "q"}

    def kind_of(c):
        if c in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(c, type) and issubclass(c, ctypes._Pointer)):
            return "P"
        return kinds.get(c, "?")

    def kind_of_decl(a):
        a = a.strip()
        if "*" in a or "ffr_stream_t" in a:
            return "P"
        t = re.sub(r"\b\w+$", "", a).replace("const ", "").strip()
        return {"int": "i", "float": "f", "int64_t": "q", "uint32_t": "I", "long long": "q"}.get(t, "?" + t)
    checked = 0
    for header in (_lib.HEADER_PATH, _lib.PROBE_HEADER_PATH):
        text = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
        for _, name, args in re.findall(r"FFR_API\s+([\w\s\*]+?)\b(ffr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
            args = args.strip()
            exp = [] if args in ("", "void") else [kind_of_decl(a) for a in args.split(",")]
            got = [kind_of(c) for c in _lib._SIGNATURES[name][1]]
            assert exp == got, (name, "".join(exp), "".join(got))
            checked += 1
    assert checked >= 60


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    rc = lib.ffr_conv3x3_bnpre_prelu_fwd(None, 1, 14, 64, None, 64, None, None, None, 0, None)
    assert rc < 0 and b"null pointer" in lib.ffr_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "conv")
    # every family of entry points validates before it touches the device: negative code + message, no exception
    one = ctypes.c_void_p(16)                                     # a non-null, never dereferenced "device pointer"
    cases = [
        (lambda: lib.ffr_conv_gemm(one, 128, 64, 64, one, 64, 64, 0, None, None, 128, 0, 0, 0, 0, 0, 0, None, None, one, 64, 0,
                           None, None, None, 0, None, 1, None, 0, 0, 0, None), b"ntaps"),
        (lambda: lib.ffr_conv_gemm(one, 128, 60, 60, one, 60, 64, 1, None, None, 128, 0, 0, 0, 0, 0, 0, None, None, one, 64, 0,
                           None, None, None, 0, None, 1, None, 0, 0, 0, None), b"multiple of 64"),
        (lambda: lib.ffr_conv_gemm(one, 128, 64, 64, one, 64, 64, 1, None, None, 128, 0, 0, 0, 0, 0, 0x1, None, None, one, 64, 0,
                           None, None, None, 0, None, 1, None, 0, 0, 0, None), b"bias"),
        (lambda: lib.ffr_conv_gemm(one, 81, 64, 64, one, 64, 64, 9, None, None, 81, 80, 9, 7, 1, 1, 0x2000, None, None, one, 64,
                           0, None, None, None, 0, None, 1, None, 0, 0, 0, None), b"pixel-major"),
        (lambda: lib.ffr_conv3x3_bnpre_prelu_fwd(one, 1, 13, 64, one, 64, one, one, one, 1, None), b"even S"),
        (lambda: lib.ffr_cosface_pack(one, 5, 70, 0, one, None, 0, None), b"rows_pad"),
        (lambda: lib.ffr_cosface_ce_fwd(one, 4, one, 300, 300, one, 30.0, 0.4, one, one, one, one, None, None), b"c_pad"),
        (lambda: lib.ffr_cosface_ce_bwd(one, 320, 300, 4, 60, one, one, one, 30.0, 0.4, one, one, None), b"bad shape"),
        (lambda: lib.ffr_wgrad(one, 60, one, 64, 0, 81, 64, 64, 9, 0, 1, 0, 64, -1, one, None, one, None), b"pitches"),
        (lambda: lib.ffr_self_similarity(one, 1, None, None, None), b"no output"),
        (lambda: lib.ffr_stem_u8_fwd(None, None, 1, one, one, one, one, 1, 112, None), b"null pointer"),
        (lambda: lib.ffr_set_conv_scratch(ctypes.c_void_p(24), 1 << 20), b"16-byte aligned"),
        (lambda: lib.ffr_set_conv_scratch(one, 512), b"flag words"),
    ]
    for call, needle in cases:
        rc = call()
        assert rc < 0 and needle in lib.ffr_last_error(), (rc, needle, lib.ffr_last_error())
    # the stream-K scratch: registering and unregistering is host-only bookkeeping; its size covers 74 CTA pairs
    assert lib.ffr_set_conv_scratch(one, lib.ffr_conv_scratch_bytes()) == 0 and lib.ffr_set_conv_scratch(None, 0) == 0
    assert lib.ffr_conv_scratch_bytes() == 0 or lib.ffr_conv_scratch_bytes() >= 1024
    assert lib.ffr_pixmajor_profitable(4) == 0 and lib.ffr_pixmajor_profitable(512) == 1
    assert lib.ffr_pixmajor_profitable(130) == 0                  # a second, nearly empty image block does not pay


def test_state_dict_layout_matches_reference():
    m = Backbone(50, 0.6, "ir_se")
    sd = ob.synth_backbone_state_dict(0)
    assert set(m.state_dict().keys()) == set(sd.keys()) and len(sd) == 402
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    m.load_state_dict(sd, strict=True)
    r = RecNet()
    rsd = orr.synth_recnet_state_dict(0)
    assert set(r.state_dict().keys()) == set(rsd.keys()) and len(rsd) == 121
    for k, v in r.state_dict().items():
        assert tuple(v.shape) == tuple(rsd[k].shape), k
    assert sum(p.numel() for p in r.parameters()) == 29925899
    assert sum(p.numel() for p in m.parameters()) == 43798720


def test_init_weights_semantics():
    torch.manual_seed(0)
    r = RecNet()
    cls_before = r.classifier.weight.detach().clone()
    init_weights(r, "kaiming")
    assert torch.equal(r.classifier.weight, cls_before)                   # AddMarginProduct keeps xavier (no 'Conv'/'Linear')
    assert float(r.Conv4Channel[0].bias.abs().max()) == 0.0               # linear biases zeroed
    w = r.Conv4Merge[0].conv2d.weight
    assert abs(float(w.std()) - (2.0 / (1536 * 9)) ** 0.5) < 2e-4          # kaiming_normal fan_in
    bn = r.Conv4Merge[0].norm.norm
    assert abs(float(bn.weight.mean()) - 1.0) < 0.01 and float(bn.bias.abs().max()) == 0.0
    assert float((r.Conv4Merge[0].relu.func.weight - 0.25).abs().max()) == 0.0


def test_modules_fail_loudly_off_gpu():
    m = Backbone(50, 0.6, "ir_se").eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 112, 112))
    with pytest.raises(RuntimeError):
        RecNet().eval()(torch.zeros(1, 512, 7, 7))
    with pytest.raises(NotImplementedError):
        Backbone(50, 0.6, "ir")


def test_layout_roundtrips():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 6, 6, generator=g)
    f = layout.to_flat(x, torch.float32)
    assert f.shape == (2 * 49, 8) and torch.equal(layout.from_flat(f, 2, 6, 8), x)
    assert float(layout.flat_pad_rows(f, 2, 6, 8).abs().max()) == 0.0
    s = layout.to_s2d(x, torch.float32)
    assert s.shape == (2 * 16, 32) and torch.equal(layout.from_s2d(s, 2, 6, 8), x)


def test_shifted_row_conv_equals_conv2d():
    """The flat-layout trick itself: a 3x3/pad-1 conv (stride 1 and, via space-to-depth, stride 2) is a sum of
    row-shifted GEMMs. Integer-valued data -> exact."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(1)
    n, c, co, S = 2, 4, 5, 6
    x = torch.randint(-3, 4, (n, c, S, S), generator=g).float()
    w = torch.randint(-2, 3, (co, c, 3, 3), generator=g).float()
    wp = w.permute(0, 2, 3, 1).reshape(co, -1)
    acc = emulate.conv_gemm(layout.to_flat(x, torch.float32), wp, c, emulate.taps_3x3_flat(S + 1))
    valid, _ = emulate.geom(n, S, x.device)
    got = layout.from_flat(acc * valid.view(-1, 1), n, S, co)
    assert torch.equal(got, F.conv2d(x, w, padding=1))
    acc2 = emulate.conv_gemm(layout.to_s2d(x, torch.float32), wp, c, emulate.taps_3x3_s2d(S // 2 + 1, c))
    valid2, _ = emulate.geom(n, S // 2, x.device)
    got2 = layout.from_flat(acc2 * valid2.view(-1, 1), n, S // 2, co)
    assert torch.equal(got2, F.conv2d(x, w, stride=2, padding=1))


def test_border_bias_table_is_exact():
    """conv(pad0(x*s+b)) == conv_{w*s}(pad0(x)) + bias9[border class] (packing.border_bias_table)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(2)
    n, c, co, S = 1, 3, 4, 5
    x = torch.randn(n, c, S, S, generator=g, dtype=torch.float64)
    w = torch.randn(co, c, 3, 3, generator=g, dtype=torch.float64)
    s = torch.rand(c, generator=g, dtype=torch.float64) + 0.5
    b = torch.randn(c, generator=g, dtype=torch.float64)
    ref = F.conv2d(x * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1), w, padding=1)
    t = torch.einsum("oirs,i->ors", w, b)
    rows = {0: (1, 2), 1: (0, 1, 2), 2: (0, 1)}
    tab = torch.stack([sum(t[:, r, q] for r in rows[ch] for q in rows[cw]) for ch in range(3) for cw in range(3)], 0)
    assert torch.allclose(tab.float(), packing.border_bias_table(w.float(), b.float()), atol=1e-5)
    _, cls = emulate.geom(n, S, x.device)
    cls_map = cls.view(n, S + 1, S + 1)[:, :S, :S]
    got = F.conv2d(x, w * s.view(1, -1, 1, 1), padding=1) + tab[cls_map].permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, atol=1e-12)


def test_packed_backbone_emulation_matches_oracle():
    """Folded/packed weights + device layouts + bf16 activation rounding (the device algorithm, emulated in torch)
    stay within the 1e-2 embedding tolerance of the fp32 oracle."""
    sd = ob.synth_backbone_state_dict(0)
    m = Backbone(50, 0.6, "ir_se").eval()
    m.load_state_dict(sd)
    pk = m._pack(torch.device("cpu"))
    x = ob.synth_faces(2, 0)
    with torch.no_grad():
        y0, f0 = ob.backbone_forward(sd, x)
        y1, f1 = emulate.backbone_forward(pk.units, (pk.stem_w, pk.stem_b, pk.stem_a), (pk.head_w, pk.head_b),
                                          (pk.bn_scale, pk.bn_shift), x)
    assert float((f0 - f1).abs().max() / f0.abs().max()) <= 1e-2
    assert float((y0 - y1).abs().max() / y0.abs().max()) <= 1.5e-2


def test_head_fold_is_exact_in_fp32():
    """pack_head: BN2d -> flatten(NCHW) -> Linear -> BN1d folded into one GEMM over the flat rows."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 512, 7, 7, generator=g)
    sd = ob.synth_backbone_state_dict(0)
    o = F.batch_norm(h, sd["output_layer.0.running_mean"], sd["output_layer.0.running_var"], sd["output_layer.0.weight"],
                     sd["output_layer.0.bias"], False, 0.0, 1e-5).reshape(2, -1)
    o = F.linear(o, sd["output_layer.3.weight"], sd["output_layer.3.bias"])
    o = F.batch_norm(o, sd["output_layer.4.running_mean"], sd["output_layer.4.running_var"], sd["output_layer.4.weight"],
                     sd["output_layer.4.bias"], False, 0.0, 1e-5)
    bn2 = packing.bn_scale_shift(sd["output_layer.0.weight"], sd["output_layer.0.bias"], sd["output_layer.0.running_mean"],
                                 sd["output_layer.0.running_var"])
    bn1 = packing.bn_scale_shift(sd["output_layer.4.weight"], sd["output_layer.4.bias"], sd["output_layer.4.running_mean"],
                                 sd["output_layer.4.running_var"])
    W = sd["output_layer.3.weight"].float().view(-1, 512, 7, 7)
    bias = bn1[0] * (sd["output_layer.3.bias"] + torch.einsum("dchw,c->d", W, bn2[1])) + bn1[1]
    Wf = (W * bn2[0].view(1, 512, 1, 1) * bn1[0].view(-1, 1, 1, 1))
    got = torch.einsum("nchw,dchw->nd", h, Wf) + bias
    assert float((got - o).abs().max()) <= 2e-4 * float(o.abs().max())
    wp, bp = packing.pack_head(sd["output_layer.3.weight"], sd["output_layer.3.bias"], bn2, bn1)
    assert wp.shape == (512, 64 * 512) and float((bp - bias).abs().max()) <= 1e-5
    flat = layout.to_flat(h, torch.float32).reshape(2, -1)
    got2 = flat @ wp.float().t() + bp
    assert float((got2 - o).abs().max()) <= 1e-2 * float(o.abs().max())


def test_chunk_bounds():
    from ffr_net_b200 import streams
    assert streams.chunk_bounds(512, 2) == [(0, 256), (256, 512)]
    assert streams.chunk_bounds(11, 3) == [(0, 11)]          # below FFR_MIN_CHUNK: one chunk
    assert streams.chunk_bounds(200, 3) == [(0, 67), (67, 134), (134, 200)]
    assert streams.chunk_bounds(0, 4) == [(0, 0)]


def test_checkpoint_container_matches_reference(tmp_path):
    """ffr_net_b200.checkpoint reads a file written by the real utils.save (golden), resolves 'latest' like
    Trainer.load_model (models/trainer.py:201-214), and round-trips the {'RecNet','optimizer','epoch','iter'} layout
    (non-strict load, optimizer state not restored)."""
    import gzip
    import os
    import shutil
    import torch
    from ffr_net_b200 import checkpoint as ck
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_ckpt_ref.pth.gzip")
    w = ck.load(gold)
    g = torch.Generator().manual_seed(11)
    assert w["epoch"] == 3 and w["iter"] == 1234 and int(w["RecNet"]["a.norm.num_batches_tracked"]) == 7
    assert torch.equal(w["RecNet"]["a.weight"], torch.randn(3, 4, generator=g))
    # save_model / load_model on a small module with the reference's container keys
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
    opt = torch.optim.Adam(net.parameters(), lr=0.1)
    d = str(tmp_path)
    p1 = ck.save_model(net, opt, d, "epoch_001", {"epoch": 1, "iter": 10})
    with gzip.open(p1, "rb") as f:                                  # a gzip stream around a torch.save payload
        assert f.read(2) in (b"PK", b"\x80\x02")
    ck.save_model(net, opt, d, "latest", {"epoch": 2, "iter": 20})
    shutil.copy(gold, os.path.join(d, "zz_other.bin"))             # not a *.pth.gzip: ignored by 'latest'
    assert ck.resolve(d, "latest").endswith("latest.pth.gzip") and ck.resolve(d, "a/b") == "a/b.pth.gzip"
    net2 = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 2))   # extra keys: strict=False
    start = ck.load_model(net2, d, "latest")
    assert start == {"epoch": 2, "iter": 20}
    assert torch.equal(net2[0].weight, net[0].weight)
    assert set(ck.load(p1).keys()) == {"RecNet", "optimizer", "epoch", "iter"}
