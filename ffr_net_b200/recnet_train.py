"""RecNet in training mode — models/recnet.py:398-429 (forward with batch-statistics BatchNorm and the label head) and
its backward, as driven by models/trainer.py:139-187. Every operation is a hand-written sm_100a kernel of
libffr_sm100 called through the C ABI; there is no autograd graph inside RecNet and no ATen / cuBLAS arithmetic:

  forward   ffr_chan_compose, ffr_recnet_prep_train (selfSimilarity, cat fan-outs, Conv4Channel, M_channel @ X, flip/cat),
            15 x [ffr_conv_gemm_ex (tcgen05 implicit GEMM, fp16 hi+lo activations x fp16 weights -> fp32 z + per-tile
            BatchNorm partial sums), ffr_bn_finalize, ffr_bn_act_fwd], ffr_feat_space_train, ffr_h9_avgpool,
            CosFace head + cross-entropy (head.py)
  backward  15 x [ffr_bn_act_bwd, ffr_wgrad (tcgen05, deterministic), ffr_conv_gemm_ex as dgrad (fp32 out)],
            ffr_feat_space_bwd, ffr_fc_bwd_gather, three tcgen05 GEMMs for the channel rectifier, ffr_chan_bwd,
            ffr_chan_compose_bwd

`TrainEngine` runs G RecNet calls (the unmasked and the masked batch of an iteration, trainer.py:144-145) as ONE
batch of G*n samples: convolutions, gradients and weight gradients see 2n rows per launch while BatchNorm statistics
(and running-statistic updates, in call order) stay per call, exactly like two sequential calls of the reference.
The public `RecNet.forward(x, label)` in training mode wraps the engine (G = 1) in a single autograd.Function, so the
reference's own Trainer.backward (`loss.backward()` over the 7-tuple) works unchanged.

Numerics: DESIGN.md "Training numerics" — fp16 hi+lo activations (~21 mantissa bits), fp16 weights, fp32 conv outputs
and activation gradients, bf16 only for dz / the weight-gradient operands; all reductions in a fixed order.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

EPI = _lib.EPI
_TAPS9 = (ctypes.c_int * 9)(*[(r - 1) * 9 + (s - 1) for r in range(3) for s in range(3)])
_FC_CHOFF = (ctypes.c_int * 3)(0, 0, 512)

# ConvLayers in execution order: (attribute path, Cin, Cout)
_CONVS = [
    ("Conv4Space.0", 561, 256), ("Conv4Space.1.conv1", 256, 256), ("Conv4Space.1.conv2", 256, 256),
    ("Conv4Space.2", 256, 128), ("Conv4Space.3.conv1", 128, 128), ("Conv4Space.3.conv2", 128, 128),
    ("Conv4Space.4", 128, 49), ("Conv4Space.5.conv1", 49, 49), ("Conv4Space.5.conv2", 49, 49),
    ("ChannelFlipMerge.0", 1024, 512), ("ChannelFlipMerge.1.conv1", 512, 512), ("ChannelFlipMerge.1.conv2", 512, 512),
    ("Conv4Merge.0", 1536, 512), ("Conv4Merge.1.conv1", 512, 512), ("Conv4Merge.1.conv2", 512, 512),
]


def _ceil64(c):
    return (c + 63) // 64 * 64


def _P(t):
    return _lib.ptr(t)


class _Act:
    """One activation matrix on the H9 grid: fp16 hi|lo [rows][2C] (forward GEMM operand) + bf16 copy [rows][C]
    (weight-gradient operand)."""

    def __init__(self, rows, c, dev, hilo=True):
        self.c = c
        self.lo = c if hilo else 0
        self.h = torch.zeros(rows, c * (2 if hilo else 1), dtype=torch.float16, device=dev)
        self.b = torch.zeros(rows, c, dtype=torch.bfloat16, device=dev)


class _Obj:
    pass


def h9_scatter_table(choff, device):
    """[81][4] (row, channel offset) table: interior row -> itself + the halo rows that mirror it (ReflectionPad2d(1))."""
    from .recnet import _h9_scatter
    return _h9_scatter(choff, device)


class TrainEngine:
    def __init__(self, model, hilo=True, deterministic=True):
        self.model = model
        self.hilo = bool(hilo)
        self.deterministic = bool(deterministic)
        self._ws = {}
        self._pack = None
        self.trace = None          # debugging: set to a list to record (name, tensor clone) after every stage
        self.layers = []
        for name, cin, cout in _CONVS:
            m = model
            for part in name.split("."):
                m = getattr(m, part) if not part.isdigit() else m[int(part)]
            L = _Obj()
            L.name, L.mod, L.cin, L.cout = name, m, cin, cout
            L.cin_p, L.cout_p = _ceil64(cin), _ceil64(cout)
            self.layers.append(L)

    def _tr(self, name, t):
        if self.trace is not None and t is not None:
            self.trace.append((name, t.detach().clone()))

    # ------------------------------------------------------------------------------------------------------
    def _packed(self, dev):
        """fp16 forward / bf16 dgrad packings of the 15 conv weights, rebuilt when the parameters change."""
        key = (_lib.weights_generation(), str(dev)) + tuple((L.mod.conv2d.weight.data_ptr(), L.mod.conv2d.weight._version)
                                                            for L in self.layers)
        if self._pack is not None and self._pack[0] == key:
            return self._pack[1]
        lib = _lib.load()
        st = _lib.stream_ptr()
        if self._pack is None or self._pack[2] != str(dev):
            bufs = [(torch.empty(L.cout_p, 9 * L.cin_p, dtype=torch.float16, device=dev),
                     torch.empty(L.cin_p, 9 * L.cout_p, dtype=torch.bfloat16, device=dev)) for L in self.layers]
        else:
            bufs = self._pack[1]                       # same addresses every step: CUDA-graph friendly
        for L, (w16, wt) in zip(self.layers, bufs):
            w = L.mod.conv2d.weight.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                raise RuntimeError("RecNet conv weights must be contiguous fp32")
            _lib.check(lib.ffr_pack_conv3x3_f16(_P(w), L.cout, L.cin, L.cout_p, L.cin_p, _P(w16), _P(wt), st),
                       "ffr_pack_conv3x3_f16")
        self._pack = (key, bufs, str(dev))
        return bufs

    def workspace(self, G, n, dev, slot=0):
        key = (G, n, str(dev), slot)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        lib = _lib.load()
        NT, R = G * n, G * n * 81
        ws = _Obj()
        ws.G, ws.n, ws.NT, ws.R = G, n, NT, R
        hl = self.hilo
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        ws.s0 = _Act(R, 576, dev, hl)
        ws.a256 = [_Act(R, 256, dev, hl) for _ in range(3)]
        ws.a128 = [_Act(R, 128, dev, hl) for _ in range(3)]
        ws.a64 = [_Act(R, 64, dev, hl) for _ in range(2)]
        ws.fm = _Act(R, 1024, dev, hl)
        ws.c512 = [_Act(R, 512, dev, hl) for _ in range(2)]
        ws.cm = _Act(R, 1536, dev, hl)
        ws.d512 = [_Act(R, 512, dev, hl) for _ in range(2)]
        ws.z = [torch.zeros(R, L.cout_p, **f32) for L in self.layers]
        ws.mr = [torch.zeros(G, 2, L.cout_p, **f32) for L in self.layers]
        pix = bool(lib.ffr_pixmajor_profitable(NT))
        ws.pix = pix
        m_tiles = 49 * ((NT + 127) // 128) if pix else (R + 127) // 128
        ws.part_rows = 4 * m_tiles
        ws.part = torch.zeros(ws.part_rows * 2 * 512, **f32)
        ws.mspace = torch.zeros(R, 64, **f32)
        ws.fs = torch.zeros(R, 512, **f32)          # feat_space (own rows)
        ws.fc = torch.zeros(R, 512, **f32)          # feat_channel after ChannelFlipMerge (own rows)
        ws.fnew = torch.zeros(R, 512, **f32)        # feat_new (own rows)
        ws.v = torch.zeros(NT, 512, **f32)
        # channel rectifier
        ws.A = torch.zeros(2112, **f32)             # A1 [0,1024) c1 [1024,1056) A2 [1056,2080) c2 [2080,2112)
        ws.g = [torch.zeros(NT * 512, 32, **f32) for _ in range(3)]
        ws.h7b = torch.zeros(NT * 512, 64, **bf)
        ws.xk = torch.zeros(NT * 512, 64, **bf)
        ws.mch2 = torch.zeros(NT * 512, 1024, dtype=torch.float16, device=dev)   # M_channel [hi | lo]
        ws.x3 = torch.zeros(NT * 64, 1536, dtype=torch.float16, device=dev)      # X^T [hi | lo | hi] per sample
        ws.fcraw = torch.zeros(NT * 512, 64, **f32)                              # feat_channel rows (c) x pixels
        ws.inv_c = torch.zeros(NT * 512, **f32)
        ws.tmat = torch.zeros(NT, 49, 32, **f32)
        # backward
        ws.dz = torch.zeros(R, 512, **bf)
        ws.da = [torch.zeros(R, 512, **f32) for _ in range(2)]
        ws.dcm = torch.zeros(R, 1024, **f32)
        ws.dfm = torch.zeros(R, 1024, **f32)
        ws.af = [torch.zeros(R, 512, **f32) for _ in range(3)]
        ws.dmsp = torch.zeros(R, 64, **f32)
        ws.bwd_rows = {c: lib.ffr_bn_act_bwd_partial_rows(G, c) for c in (64, 128, 256, 512)}
        ws.bwd_part = torch.zeros(max(r * 3 * c for c, r in ws.bwd_rows.items()), **f32)
        ws.gsum = torch.zeros(G * 2 * 512, **f32)
        wg = 0
        for L in self.layers:
            wg = max(wg, lib.ffr_wgrad_workspace_floats(R, L.cout, L.cin, 9, 1 if self.deterministic else 0))
        wg = max(wg, lib.ffr_wgrad_workspace_floats(NT * 512, 512, 33, 1, 1 if self.deterministic else 0))
        ws.wgrad_ws = torch.zeros(int(wg), **f32)
        ws.dfc_op = torch.zeros(NT * 512, 64, **bf)
        ws.dmpre = torch.zeros(NT * 512, 512, **bf)
        ws.dh7 = torch.zeros(NT * 512, 64, **f32)
        ws.w8t = torch.zeros(64, 512, **bf)
        ws.chan_part = torch.zeros(NT * lib.ffr_chan_bwd_part_floats(), **f32)
        ws.dslope_part = torch.zeros(NT * 3 * 512, **f32)
        ws.chan_tmp = torch.zeros(2112, **f32)
        ws.t_h9 = {off: h9_scatter_table(off, dev) for off in (0, 512)}
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------------------
    def _conv_fwd(self, lib, ws, L, w16, src, z, st):
        d = _lib.ConvGemmDesc()
        d.a, d.a_rows, d.a_cols, d.a_ld = _P(src.h), ws.R, src.h.shape[1], src.h.shape[1]
        d.wp, d.Cin, d.Cout, d.ntaps = _P(w16), L.cin_p, L.cout_p, 9
        d.tap_row_shift = ctypes.cast(_TAPS9, ctypes.c_void_p)
        d.M, d.rows_per_img, d.Wp, d.S, d.h0, d.n_img = ws.R, 81, 9, 7, 1, ws.NT
        d.flags = EPI.GEOM | EPI.STATS | EPI.OUT_F32 | (EPI.PIXMAJOR if ws.pix else 0)
        d.out_f32 = _P(z)
        d.stats_part = _P(ws.part)
        d.num_splits = 1
        d.a_hilo, d.a_lo_off, d.f16 = (1 if src.lo else 0), src.lo, 1
        _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "conv fwd " + L.name)

    def _layer_fwd(self, lib, ws, i, packs, src, dst, st, res=None, table=None, out_f=None, sigmoid=False):
        """ConvLayer.forward (recnet.py:78-85) [+ residual add, :217] for the whole batch."""
        L = self.layers[i]
        bn = L.mod.norm.norm
        z, mr = ws.z[i], ws.mr[i]
        self._conv_fwd(lib, ws, L, packs[i][0], src, z, st)
        self._tr("fwd.%s.z" % L.name, z)
        self._tr("fwd.%s.part" % L.name, ws.part[: ws.part_rows * 2 * L.cout_p])
        _lib.check(lib.ffr_bn_finalize(_P(ws.part), ws.part_rows, 1 if ws.pix else 0, ws.NT, ws.n, L.cout_p, L.cout,
                                       BN_MOMENTUM, BN_EPS, _P(bn.running_mean), _P(bn.running_var),
                                       _P(bn.num_batches_tracked), _P(mr), st), "bn_finalize " + L.name)
        tab = table if table is not None else ws.t_h9[0]
        _lib.check(lib.ffr_bn_act_fwd(
            _P(z), L.cout_p, _P(mr), _P(bn.weight), _P(bn.bias), _P(L.mod.relu.func.weight),
            _P(res.h) if res is not None else None, res.h.shape[1] if res is not None else 0,
            res.lo if res is not None else 0,
            _P(dst.h) if dst is not None else None, dst.h.shape[1] if dst is not None else 0,
            dst.lo if dst is not None else 0,
            _P(dst.b) if dst is not None else None, dst.b.shape[1] if dst is not None else 0,
            _P(out_f), out_f.shape[1] if out_f is not None else 0, 1 if sigmoid else 0,
            _P(tab), 4, ws.NT, ws.n, L.cout_p, L.cout, st), "bn_act_fwd " + L.name)
        self._tr("fwd.%s.mr" % L.name, mr)
        if dst is not None:
            self._tr("fwd.%s.out_h" % L.name, dst.h)
            self._tr("fwd.%s.out_b" % L.name, dst.b)
        self._tr("fwd.%s.out_f" % L.name, out_f)

    def forward(self, x, n_groups, slot=0, v_out=None):
        """x: (G*n,512,7,7) fp32 CUDA, the G calls concatenated. Returns the workspace holding every result:
        ws.v (pooled feat_new, (G*n,512); written to v_out instead when given), ws.fs / ws.fc / ws.fnew (fp32 H9 rows),
        ws.mspace, ws.mch2. `slot` selects one of several resident workspaces of the same shape."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        m = self.model
        dev = x.device
        NT = x.shape[0]
        if NT % n_groups:
            raise ValueError("batch %d is not %d equal groups" % (NT, n_groups))
        n = NT // n_groups
        if n_groups > 1 and n % 32:
            raise ValueError("batched RecNet calls need a per-call batch that is a multiple of 32 (got %d)" % n)
        for bn in (L.mod.norm.norm for L in self.layers):
            if bn.num_batches_tracked.dtype != torch.int64:
                raise RuntimeError("num_batches_tracked must be int64")
        ws = self.workspace(n_groups, n, dev, slot)
        ws.x = x
        packs = self._packed(dev)
        c = m.Conv4Channel
        A = ws.A
        _lib.check(lib.ffr_chan_compose(_P(c[2].weight), _P(c[2].bias), _P(c[3].weight), _P(c[3].bias),
                                        _P(c[5].weight), _P(c[5].bias), _P(c[6].weight), _P(c[6].bias),
                                        _P(A[0:]), _P(A[1024:]), _P(A[1056:]), _P(A[2080:]), _P(c[8].weight), _P(ws.w8t), st),
                   "chan_compose")
        d = _lib.PrepTrainDesc()
        d.x = _P(x)
        d.w0, d.b0 = _P(c[0].weight), _P(c[0].bias)
        d.slope1, d.slope4, d.slope7 = _P(c[1].func.weight), _P(c[4].func.weight), _P(c[7].func.weight)
        d.A1, d.c1, d.A2, d.c2 = _P(A[0:]), _P(A[1024:]), _P(A[1056:]), _P(A[2080:])
        d.w8, d.b8 = _P(c[8].weight), _P(c[8].bias)
        for nm, act in (("s0", ws.s0), ("cm", ws.cm)):
            setattr(d, nm + "_h", _P(act.h)); setattr(d, nm + "_ld", act.h.shape[1]); setattr(d, nm + "_lo", act.lo)
            setattr(d, nm + "_b", _P(act.b)); setattr(d, nm + "_ldb", act.b.shape[1])
        d.g0, d.g1, d.g2 = _P(ws.g[0]), _P(ws.g[1]), _P(ws.g[2])
        d.h7b, d.xk, d.mch2, d.x3 = _P(ws.h7b), _P(ws.xk), _P(ws.mch2), _P(ws.x3)
        d.inv_c, d.tmat, d.ss_space = _P(ws.inv_c), _P(ws.tmat), None
        _lib.check(lib.ffr_recnet_prep_train(ctypes.byref(d), NT, st), "recnet_prep_train")
        # feat_channel = M_channel @ X (recnet.py:410): rows (sample, c) x pixels; three K = 512 "taps" over A = [hi | lo]
        # (column offsets 0, 0, 512) against B = [X_hi | X_lo | X_hi]: hi.hi + hi.lo + lo.hi in fp16
        g = _lib.ConvGemmDesc()
        g.a, g.a_rows, g.a_cols, g.a_ld = _P(ws.mch2), NT * 512, 1024, 1024
        g.wp, g.Cin, g.Cout, g.ntaps = _P(ws.x3), 512, 64, 3
        g.tap_ch_off = ctypes.cast(_FC_CHOFF, ctypes.c_void_p)
        g.M = NT * 512
        g.flags = EPI.OUT_F32
        g.out_f32 = _P(ws.fcraw)
        g.num_splits, g.b_rows_per_mtile, g.b_mtile_div, g.f16 = 1, 64, 4, 1
        _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(g), st), "feat_channel GEMM")
        _lib.check(lib.ffr_fc_scatter(_P(ws.fcraw), _P(ws.fm.h), ws.fm.h.shape[1], ws.fm.lo, _P(ws.fm.b), ws.fm.b.shape[1],
                                      NT, st), "fc_scatter")                      # flip / cat fan-out (:416-417)
        for nm in ("s0", "cm", "fm"):
            self._tr("fwd.prep.%s" % nm, getattr(ws, nm).h)
        self._tr("fwd.prep.g2", ws.g[2])
        self._tr("fwd.prep.mch2", ws.mch2)

        f = lambda *a, **k: self._layer_fwd(lib, ws, *a, st=st, **k)
        # spatial rectifier (recnet.py:362-371, :404-405)
        f(0, packs, ws.s0, ws.a256[0])
        f(1, packs, ws.a256[0], ws.a256[1])
        f(2, packs, ws.a256[1], ws.a256[2], res=ws.a256[0])
        f(3, packs, ws.a256[2], ws.a128[0])
        f(4, packs, ws.a128[0], ws.a128[1])
        f(5, packs, ws.a128[1], ws.a128[2], res=ws.a128[0])
        f(6, packs, ws.a128[2], ws.a64[0])
        f(7, packs, ws.a64[0], ws.a64[1])
        f(8, packs, ws.a64[1], None, res=ws.a64[0], out_f=ws.mspace, sigmoid=True)
        _lib.check(lib.ffr_feat_space_train(_P(x), _P(ws.mspace), _P(ws.cm.h), ws.cm.h.shape[1], ws.cm.lo, _P(ws.cm.b),
                                            ws.cm.b.shape[1], _P(ws.fs), 512, NT, st), "feat_space_train")
        # flip merge (:415-418) and final merge (:420-421)
        f(9, packs, ws.fm, ws.c512[0])
        f(10, packs, ws.c512[0], ws.c512[1])
        f(11, packs, ws.c512[1], ws.cm, res=ws.c512[0], table=ws.t_h9[512], out_f=ws.fc)
        f(12, packs, ws.cm, ws.d512[0])
        f(13, packs, ws.d512[0], ws.d512[1])
        f(14, packs, ws.d512[1], None, res=ws.d512[0], out_f=ws.fnew)
        ws.v_cur = v_out if v_out is not None else ws.v
        _lib.check(lib.ffr_h9_avgpool(_P(ws.fnew), 512, _P(ws.v_cur), 512, NT, 512, st), "h9_avgpool")   # :423
        return ws

    # ------------------------------------------------------------------------------------------------------
    def _dgrad(self, lib, ws, L, wt, dz, out, cout_used, st):
        d = _lib.ConvGemmDesc()
        d.a, d.a_rows, d.a_cols, d.a_ld = _P(dz), ws.R, L.cout_p, dz.shape[1]
        d.wp, d.Cin, d.Cout, d.ntaps = _P(wt), L.cout_p, cout_used, 9
        d.tap_row_shift = ctypes.cast(_TAPS9, ctypes.c_void_p)
        d.M, d.rows_per_img, d.Wp, d.S, d.h0, d.n_img = ws.R, 81, 9, 7, 1, ws.NT
        d.flags = EPI.OUT_F32 | ((EPI.PIXMAJOR | EPI.PIX_DGRAD) if ws.pix else 0)
        d.out_f32 = _P(out)
        d.num_splits = 1
        _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dgrad " + L.name)

    def _layer_bwd(self, lib, ws, i, packs, grads, src, st, da=None, da_ch0=0, dadd=None, dadd_ch0=0, dv=None,
                   afold=None, dgrad_out=None, dgrad_cols=None, accumulate=False):
        """Backward of one ConvLayer: BatchNorm / PReLU backward, weight gradient, data gradient."""
        L = self.layers[i]
        bn = L.mod.norm.norm
        C = L.cout_p
        dz = ws.dz[:, :C] if C == 512 else ws.dz.view(-1)[: ws.R * C].view(ws.R, C)
        af = afold.view(-1)[: ws.R * C].view(ws.R, C)
        gw, gb, gs = grads[L.name + ".norm.norm.weight"], grads[L.name + ".norm.norm.bias"], grads[L.name + ".relu.func.weight"]
        _lib.check(lib.ffr_bn_act_bwd(
            _P(da), da.shape[1] if da is not None else 0, da_ch0, _P(ws.t_h9[0]), 4,
            _P(dadd), dadd.shape[1] if dadd is not None else 0, dadd_ch0,
            _P(dv), 512 if dv is not None else 0, 1.0 / 49.0,
            _P(ws.z[i]), C, _P(ws.mr[i]), _P(bn.weight), _P(bn.bias), _P(L.mod.relu.func.weight),
            _P(af), C, _P(ws.bwd_part), _P(ws.gsum), _P(gw), _P(gb), _P(gs), 1 if accumulate else 0, L.cout,
            _P(dz), C, ws.NT, ws.n, C, st), "bn_act_bwd " + L.name)
        self._tr("bwd.%s.afold" % L.name, af)
        self._tr("bwd.%s.dz" % L.name, dz)
        self._tr("bwd.%s.dgamma" % L.name, gw)
        _lib.check(lib.ffr_wgrad(_P(dz), C, _P(src.b), src.b.shape[1], 0, ws.R, L.cout, L.cin, 9, 0,
                                 1 if self.deterministic else 0, 1 if accumulate else 0, L.cin, -1,
                                 _P(grads[L.name + ".conv2d.weight"]), None, _P(ws.wgrad_ws), st), "wgrad " + L.name)
        self._tr("bwd.%s.dw" % L.name, grads[L.name + ".conv2d.weight"])
        if dgrad_out is not None:
            self._dgrad(lib, ws, L, packs[i][1], dz, dgrad_out, dgrad_cols if dgrad_cols is not None else L.cin_p, st)
            self._tr("bwd.%s.dgrad" % L.name, dgrad_out)
        return af

    def backward(self, ws, grads, dv=None, dfs=None, dfc=None, dfnew=None, accumulate=False, on_stage=None):
        """Gradients of every RecNet parameter except the classifier, written to `grads[name]` (fp32 tensors shaped like
        the parameters; overwritten unless accumulate). dv: (G*n,512) gradient of the pooled feature; dfs / dfc / dfnew:
        optional fp32 H9 matrices [R][512] (own rows) with the gradients w.r.t. feat_space / feat_channel / feat_new.
        on_stage(prefix) is called when every gradient of the parameters under `prefix` has been written (the trainer
        launches that bucket's all-reduce)."""
        stage = on_stage if on_stage is not None else (lambda prefix: None)
        lib = _lib.load()
        st = _lib.stream_ptr()
        packs = self._packed(ws.x.device)
        R = ws.R
        view = lambda t, c: t.view(-1)[: R * c].view(R, c)
        b = lambda i, src, **k: self._layer_bwd(lib, ws, i, packs, grads, src, st, accumulate=accumulate, **k)
        # ---- Conv4Merge (layers 12..14) ----
        af14 = b(14, ws.d512[1], dv=dv, dadd=dfnew, afold=ws.af[0], dgrad_out=ws.da[0])
        b(13, ws.d512[0], da=ws.da[0], afold=ws.af[1], dgrad_out=ws.da[1])
        b(12, ws.cm, da=ws.da[1], dadd=af14, afold=ws.af[2], dgrad_out=ws.dcm, dgrad_cols=1024)
        stage("Conv4Merge")
        # ---- ChannelFlipMerge (9..11): its output sits in slot [512,1024) of the Conv4Merge input ----
        af11 = b(11, ws.c512[1], da=ws.dcm, da_ch0=512, dadd=dfc, afold=ws.af[0], dgrad_out=ws.da[0])
        b(10, ws.c512[0], da=ws.da[0], afold=ws.af[1], dgrad_out=ws.da[1])
        b(9, ws.fm, da=ws.da[1], dadd=af11, afold=ws.af[2], dgrad_out=ws.dfm)
        stage("ChannelFlipMerge")
        # ---- channel rectifier: feat_channel = M_channel @ X, M_channel = sigmoid(h7 W8^T + b8) ----
        m = self.model
        c = m.Conv4Channel
        NT = ws.NT
        _lib.check(lib.ffr_fc_bwd_gather(_P(ws.dfm), 1024, _P(ws.dfc_op), NT, st), "fc_bwd_gather")
        d = _lib.ConvGemmDesc()            # dM_pre[c][j] = (sum_hw dFC[c][hw] X[j][hw]) * m (1 - m)
        d.a, d.a_rows, d.a_cols, d.a_ld = _P(ws.dfc_op), NT * 512, 64, 64
        d.wp, d.Cin, d.Cout, d.ntaps = _P(ws.xk), 64, 512, 1
        d.M = NT * 512
        d.flags = EPI.MUL_DSIG | EPI.RES_F16
        d.out, d.ldo = _P(ws.dmpre), 512
        d.res, d.ldres = _P(ws.mch2), 1024         # m = hi part of the fp16 M_channel
        d.num_splits, d.b_rows_per_mtile, d.b_mtile_div = 1, 512, 4
        _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dM_pre GEMM")
        self._tr("bwd.chan.dfc_op", ws.dfc_op)
        self._tr("bwd.chan.dmpre", ws.dmpre)
        d = _lib.ConvGemmDesc()            # dh7[c][k] = sum_j dM_pre[c][j] W8[j][k]  (W8^T packed by ffr_chan_compose)
        d.a, d.a_rows, d.a_cols, d.a_ld = _P(ws.dmpre), NT * 512, 512, 512
        d.wp, d.Cin, d.Cout, d.ntaps = _P(ws.w8t), 512, 64, 1
        d.M = NT * 512
        d.flags = EPI.OUT_F32
        d.out_f32 = _P(ws.dh7)
        d.num_splits = 1
        _lib.check(lib.ffr_conv_gemm_ex(ctypes.byref(d), st), "dh7 GEMM")
        self._tr("bwd.chan.dh7", ws.dh7)
        # dW8[j][k] = sum_rows dM_pre[row][j] h7[row][k]; the ones column of h7b gives db8[j]
        _lib.check(lib.ffr_wgrad(_P(ws.dmpre), 512, _P(ws.h7b), 64, 0, NT * 512, 512, 33, 1, 0,
                                 1 if self.deterministic else 0, 1 if accumulate else 0, 32, 32,
                                 _P(grads["Conv4Channel.8.weight"]), _P(grads["Conv4Channel.8.bias"]),
                                 _P(ws.wgrad_ws), st), "wgrad W8")
        A = ws.A
        _lib.check(lib.ffr_chan_bwd(_P(ws.x), _P(ws.dh7), _P(ws.g[0]), _P(ws.g[1]), _P(ws.g[2]), _P(ws.inv_c),
                                    _P(ws.tmat), _P(A[0:]), _P(A[1056:]), _P(c[1].func.weight), _P(c[4].func.weight),
                                    _P(c[7].func.weight), _P(ws.chan_part), _P(ws.dslope_part), _P(ws.chan_tmp),
                                    _P(grads["Conv4Channel.0.bias"]), _P(grads["Conv4Channel.0.weight"]),
                                    _P(grads["Conv4Channel.1.func.weight"]), _P(grads["Conv4Channel.4.func.weight"]),
                                    _P(grads["Conv4Channel.7.func.weight"]), 1 if accumulate else 0, NT, st), "chan_bwd")
        T = ws.chan_tmp                    # dA2 [0,1024) dc2 [1024,1056) dA1 [1056,2080) dc1 [2080,2112)
        _lib.check(lib.ffr_chan_compose_bwd(
            _P(c[2].weight), _P(c[2].bias), _P(c[3].weight), _P(c[5].weight), _P(c[5].bias), _P(c[6].weight),
            _P(T[1056:]), _P(T[2080:]), _P(T[0:]), _P(T[1024:]),
            _P(grads["Conv4Channel.2.weight"]), _P(grads["Conv4Channel.2.bias"]), _P(grads["Conv4Channel.3.weight"]),
            _P(grads["Conv4Channel.3.bias"]), _P(grads["Conv4Channel.5.weight"]), _P(grads["Conv4Channel.5.bias"]),
            _P(grads["Conv4Channel.6.weight"]), _P(grads["Conv4Channel.6.bias"]), 1 if accumulate else 0, st),
            "chan_compose_bwd")
        stage("Conv4Channel")
        # ---- spatial rectifier: feat_space = X @ M_space (slot [0,512) of the Conv4Merge input) ----
        _lib.check(lib.ffr_feat_space_bwd(_P(ws.x), _P(ws.mspace), _P(ws.dcm), 1024, _P(dfs), 512 if dfs is not None else 0,
                                          _P(ws.dmsp), NT, st), "feat_space_bwd")
        self._tr("bwd.space.dmsp", ws.dmsp)
        af8 = b(8, ws.a64[1], dadd=ws.dmsp, afold=ws.af[0], dgrad_out=view(ws.da[0], 64))
        b(7, ws.a64[0], da=view(ws.da[0], 64), afold=ws.af[1], dgrad_out=view(ws.da[1], 64))
        b(6, ws.a128[2], da=view(ws.da[1], 64), dadd=af8, afold=ws.af[2], dgrad_out=view(ws.da[0], 128))
        af5 = b(5, ws.a128[1], da=view(ws.da[0], 128), afold=ws.af[0], dgrad_out=view(ws.da[1], 128))
        b(4, ws.a128[0], da=view(ws.da[1], 128), afold=ws.af[1], dgrad_out=view(ws.da[0], 128))
        b(3, ws.a256[2], da=view(ws.da[0], 128), dadd=af5, afold=ws.af[2], dgrad_out=view(ws.da[1], 256))
        af2 = b(2, ws.a256[1], da=view(ws.da[1], 256), afold=ws.af[0], dgrad_out=view(ws.da[0], 256))
        b(1, ws.a256[0], da=view(ws.da[0], 256), afold=ws.af[1], dgrad_out=view(ws.da[1], 256))
        b(0, ws.s0, da=view(ws.da[1], 256), dadd=af2, afold=ws.af[2])
        stage("Conv4Space")


def engine(model):
    eng = getattr(model, "_train_engine", None)
    if eng is None:
        eng = TrainEngine(model)
        model._train_engine = eng
    return eng


def param_grad_names(model):
    return [k for k, _ in model.named_parameters()]


# ----------------------------------------------------------------------------------------------------------
# Public training-mode forward: one autograd.Function around the engine (G = 1).
# ----------------------------------------------------------------------------------------------------------
def _rows_to_nchw(lib, rows, ld, c, n, st, ch0=0):
    y = torch.empty(n, c, 7, 7, dtype=torch.float32, device=rows.device)
    _lib.check(lib.ffr_rows_to_nchw(_P(rows), 1, ld, ch0, None, None, _P(y), n, 7, 9, 1, 81, c, st), "rows_to_nchw")
    return y


class _RecNetTrainFn(torch.autograd.Function):
    """(x, *parameters) -> (feat_new_v, feat_new, feat_space, feat_channel, M_space, M_channel). The gradient flows to the
    75 RecNet parameters below the classifier; M_space / M_channel are returned without gradient (the reference's
    trainer never differentiates through them, models/trainer.py:154-180), x is the frozen backbone's output."""

    @staticmethod
    def forward(ctx, model, x, *params):
        lib = _lib.load()
        st = _lib.stream_ptr()
        eng = engine(model)
        n = x.shape[0]
        # the reference's trainer keeps two forward calls alive until loss.backward() (trainer.py:144-145, :180): the
        # calls rotate over `train_slots` resident workspaces
        slots = int(getattr(model, "train_slots", 2))
        eng._fwd_count = getattr(eng, "_fwd_count", 0) + 1
        slot = eng._fwd_count % slots
        ws = eng.forward(x, 1, slot=slot)
        ws.owner = eng._fwd_count
        v = ws.v_cur.clone()
        fnew = _rows_to_nchw(lib, ws.fnew, 512, 512, n, st)
        fs = _rows_to_nchw(lib, ws.fs, 512, 512, n, st)
        fc = _rows_to_nchw(lib, ws.fc, 512, 512, n, st)
        msp = _rows_to_nchw(lib, ws.mspace, 64, 64, n, st)[:, :49].reshape(n, 49, 49).contiguous()
        mch = ws.mch2[:, :512].float().view(n, 512, 512)
        ctx.model, ctx.ws, ctx.n = model, ws, n
        ctx.step = ws.owner
        ctx.mark_non_differentiable(msp, mch)
        ctx.set_materialize_grads(False)
        return v, fnew, fs, fc, msp, mch

    @staticmethod
    def backward(ctx, gv, gfnew, gfs, gfc, _gmsp, _gmch):
        lib = _lib.load()
        st = _lib.stream_ptr()
        model, ws, n = ctx.model, ctx.ws, ctx.n
        eng = engine(model)
        if ws.owner != ctx.step:
            raise RuntimeError("RecNet training workspace was overwritten by a later forward call before backward(): more "
                               "than model.train_slots (= %d) forward calls were alive; raise model.train_slots"
                               % int(getattr(model, "train_slots", 2)))
        dev = ws.x.device

        def h9(g):
            if g is None:
                return None
            buf = torch.zeros(ws.R, 512, dtype=torch.float32, device=dev)
            _lib.check(lib.ffr_nchw_to_h9_f32(_P(g.contiguous().float()), _P(buf), 512, 0, n, 512, st), "nchw_to_h9_f32")
            return buf
        names = [k for k, _ in model.named_parameters() if k != "classifier.weight"]
        grads = {k: torch.zeros_like(p) for k, p in model.named_parameters() if k != "classifier.weight"}
        dv = gv.contiguous().float() if gv is not None else torch.zeros(n, 512, dtype=torch.float32, device=dev)
        eng.backward(ws, grads, dv=dv, dfs=h9(gfs), dfc=h9(gfc), dfnew=h9(gfnew))
        return (None, None) + tuple(grads[k] for k in names)


def cosine_sim(x1, x2, dim=1):
    """recnet.py:220-224."""
    x1 = F.normalize(x1, dim=2)
    x2 = F.normalize(x2, dim=2)
    return torch.bmm(x1, x2.permute(0, 2, 1))


def self_similarity(x):
    """selfSimilarity, recnet.py:226-236. Without autograd (the no-grad targets of the loss, trainer.py:157, or any
    inference use) both Grams come from the library's fp32 kernels. With a gradient (the reference's own
    Trainer.backward calling this helper) the literal op sequence runs under autograd; this repo's Trainer does not
    use it — its similarity losses are the fused kernels of csrc/loss_kernels.cu."""
    if not x.is_cuda:
        raise RuntimeError("ffr_net_b200.selfSimilarity runs only on CUDA")
    if not (torch.is_grad_enabled() and x.requires_grad) and tuple(x.shape[1:]) == (512, 7, 7):
        lib = _lib.load()
        xc = x.detach().contiguous().float()
        n = xc.shape[0]
        ss_s = torch.empty(n, 49, 7, 7, dtype=torch.float32, device=x.device)
        ss_c = torch.empty(n, 512, 512, dtype=torch.float32, device=x.device)
        _lib.check(lib.ffr_self_similarity(_lib.ptr(xc), n, _lib.ptr(ss_s), _lib.ptr(ss_c), _lib.stream_ptr()),
                   "ffr_self_similarity")
        return ss_s, ss_c
    h, w = x.size(2), x.size(3)
    v = x.reshape(x.size(0), x.size(1), -1)
    ss_space = cosine_sim(v.permute(0, 2, 1), v.permute(0, 2, 1))
    ss_channel = cosine_sim(v, v)
    return ss_space.reshape(ss_space.size(0), ss_space.size(1), h, w), ss_channel


def add_margin_product(weight, x, label, s=30.0, m=0.40):
    """AddMarginProduct.forward, recnet.py:257-270 — the literal (N,10575) outputs of the public 7-tuple (one-hot built
    on the input's device). The Trainer uses the fused head (head.py) instead."""
    cosine = F.linear(F.normalize(x), F.normalize(weight))
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    output = (one_hot * (cosine - m)) + ((1.0 - one_hot) * cosine)
    return output * s, cosine


def forward_train(model, x, label):
    """RecNet.forward in training mode (recnet.py:398-429): 2-tuple without label, the reference's 7-tuple with it."""
    if not x.is_cuda:
        raise RuntimeError("ffr_net_b200.RecNet runs only on CUDA (sm_100a); there is no CPU fallback")
    if x.requires_grad:
        raise NotImplementedError("RecNet's input is the frozen backbone's feature map (models/trainer.py:62-63, :141-142): "
                                  "a gradient w.r.t. it is not implemented")
    x = x.detach().contiguous().float()
    params = [p for k, p in model.named_parameters() if k != "classifier.weight"]
    v, fnew, fs, fc, msp, mch = _RecNetTrainFn.apply(model, x, *params)
    if label is None:
        return v, fnew                                                                     # :426
    pred_loss, pred_label = add_margin_product(model.classifier.weight, v, label, model.classifier.s,
                                               model.classifier.m)                        # :428
    return v, pred_loss, pred_label, msp, mch, fs, fc
