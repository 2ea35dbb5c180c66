"""RecNet (feature rectification) — drop-in for `models/recnet.py` of the reference.

Same class names, constructor arguments, forward signature and state_dict key layout (121 keys) as
/root/reference/models/recnet.py:52-143,202-277,342-429. The forward pass runs in the sm_100a library: one prep
kernel (self-similarity, concat fan-out, thin channel-rectifier chain), the tcgen05 implicit-GEMM kernel for the 15
reflection-padded 3x3 convolutions (BatchNorm folded, PReLU / residual / sigmoid fused, concatenations and the W-flip
expressed as scatter tables) and for the two per-sample GEMMs of the channel rectifier.

The torch sub-modules only hold parameters under the reference's names. There is no CPU / PyTorch fallback.
"""
import ctypes
import os
import math

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init

from . import _lib, packing, streams

READY = True   # bench.py includes the RecBlock stage when this module is importable and READY

NUM_CLASSES = 10575


def init_weights(net, init_type="normal", init_gain=0.02):
    """Reference recnet.py:13-42: conv/linear weights by `init_type`, biases 0, BatchNorm2d weight ~ N(1, gain)."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)


class ReluLayer(nn.Module):
    """Parameter holder (reference :87-115); only 'prelu' and 'none' are used by RecNet."""

    def __init__(self, channels, relu_type="relu"):
        super().__init__()
        relu_type = relu_type.lower()
        if relu_type == "prelu":
            self.func = nn.PReLU(channels)
        elif relu_type == "none":
            self.func = None
        else:
            raise NotImplementedError("relu type %s is not used by FFR-Net's RecNet" % relu_type)


class NormLayer(nn.Module):
    """Parameter holder (reference :117-143); only 'bn' and 'none' are used by RecNet."""

    def __init__(self, channels, norm_type="bn"):
        super().__init__()
        norm_type = norm_type.lower()
        if norm_type == "bn":
            self.norm = nn.BatchNorm2d(channels)
        elif norm_type == "none":
            self.norm = None
        else:
            raise NotImplementedError("norm type %s is not used by FFR-Net's RecNet" % norm_type)


class ConvLayer(nn.Module):
    """ReflectionPad2d(1) -> Conv2d 3x3 (bias only without norm) -> norm -> relu (reference :52-85)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, scale="none", norm_type="none", relu_type="none",
                 use_pad=True, use_sn=False, groups=1):
        super().__init__()
        if scale != "none" or use_sn or groups != 1 or kernel_size != 3 or not use_pad:
            raise NotImplementedError("only the ConvLayer configuration RecNet uses is implemented")
        bias = norm_type in ["pixel", "none"]
        self.reflection_pad = nn.ReflectionPad2d(kernel_size // 2)
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, 1, bias=bias, groups=groups)
        self.relu = ReluLayer(out_channels, relu_type)
        self.norm = NormLayer(out_channels, norm_type=norm_type)


class ResidualBlock(nn.Module):
    def __init__(self, inplanes, planes, kernel_size=3, norm_type="none", relu_type="none"):
        super().__init__()
        conv_args = {"norm_type": norm_type, "relu_type": relu_type}
        self.conv1 = ConvLayer(inplanes, planes, kernel_size, **conv_args)
        self.conv2 = ConvLayer(planes, planes, kernel_size, **conv_args)


class AddMarginProduct(nn.Module):
    """CosFace head parameter holder (reference :238-277)."""

    def __init__(self, in_features, out_features=NUM_CLASSES, s=30.0, m=0.40):
        super().__init__()
        self.in_features, self.out_features, self.s, self.m = in_features, out_features, s, m
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        nn.init.xavier_uniform_(self.weight)

    def __repr__(self):
        return "%s(in_features=%d, out_features=%d, s=%s, m=%s)" % (
            self.__class__.__name__, self.in_features, self.out_features, self.s, self.m)


def l2_norm(input, axis=1):
    return input / torch.norm(input, 2, axis, True)


class _Obj:
    pass


def _ceil64(c):
    return (c + 63) // 64 * 64


def _h9_scatter(choff, device):
    """Scatter table of the H9 grid: every interior row goes to itself and to the halo rows that mirror it
    (ReflectionPad2d(1): padded index 0 <- interior 1, padded index 8 <- interior 5)."""
    t = torch.full((81, 4, 2), -1, dtype=torch.int32)
    for hp in range(1, 8):
        for wp in range(1, 8):
            h, w = hp - 1, wp - 1
            mh = -2 if h == 1 else (2 if h == 5 else 0)
            mw = -2 if w == 1 else (2 if w == 5 else 0)
            dst = [(hp, wp)]
            if mh:
                dst.append((hp + mh, wp))
            if mw:
                dst.append((hp, wp + mw))
            if mh and mw:
                dst.append((hp + mh, wp + mw))
            for k, (a, b) in enumerate(dst):
                t[hp * 9 + wp, k, 0] = a * 9 + b
                t[hp * 9 + wp, k, 1] = choff
    return t.to(device).contiguous()


def _flip_scatter(device):
    """Rows (h*7+w) of feat_channel -> ChannelFlipMerge input (recnet.py:416-417): slot [512,1024) gets the map itself,
    slot [0,512) the W-flipped map; both with their reflection mirrors. 128 rows per sample, 49 valid."""
    t = torch.full((128, 8, 2), -1, dtype=torch.int32)
    for h in range(7):
        for w in range(7):
            k = 0
            for ww, ch in ((w, 512), (6 - w, 0)):
                mh = -2 if h == 1 else (2 if h == 5 else 0)
                mw = -2 if ww == 1 else (2 if ww == 5 else 0)
                hp, wp = h + 1, ww + 1
                dst = [(hp, wp), (hp + mh, wp) if mh else None, (hp, wp + mw) if mw else None,
                       (hp + mh, wp + mw) if (mh and mw) else None]
                for d in dst:
                    if d is not None:
                        t[h * 7 + w, k, 0] = d[0] * 9 + d[1]
                        t[h * 7 + w, k, 1] = ch
                    k += 1
            # compact: valid entries first (entry 0 must be valid for the row to count as valid)
            ent = [tuple(e.tolist()) for e in t[h * 7 + w] if e[0] >= 0]
            t[h * 7 + w] = -1
            for i, e in enumerate(ent):
                t[h * 7 + w, i, 0], t[h * 7 + w, i, 1] = e
    return t.to(device).contiguous()


class RecNet(nn.Module):
    def __init__(self, channel=512, shape=7, norm_type="bn", relu_type="prelu"):
        super().__init__()
        if channel != 512 or shape != 7 or norm_type != "bn" or relu_type != "prelu":
            raise NotImplementedError("the CUDA path implements RecNet(512, 7, 'bn', 'prelu') (models/trainer.py:60)")
        self.channel, self.shape = channel, shape
        ca = {"norm_type": norm_type, "relu_type": relu_type}
        hw = shape ** 2
        self.Conv4Space = nn.Sequential(
            ConvLayer(channel + hw, 256, **ca), ResidualBlock(256, 256, **ca),
            ConvLayer(256, 128, **ca), ResidualBlock(128, 128, **ca),
            ConvLayer(128, hw, **ca), ResidualBlock(hw, hw, **ca),
            nn.Sigmoid())
        self.Conv4Channel = nn.Sequential(
            nn.Linear(channel + hw, 32), ReluLayer(512, "prelu"), nn.Linear(32, channel),
            nn.Linear(channel, 32), ReluLayer(512, "prelu"), nn.Linear(32, channel),
            nn.Linear(channel, 32), ReluLayer(512, "prelu"), nn.Linear(32, channel),
            nn.Sigmoid())
        self.ChannelFlipMerge = nn.Sequential(ConvLayer(channel * 2, channel, **ca), ResidualBlock(channel, channel, **ca))
        self.Conv4Merge = nn.Sequential(ConvLayer(channel * 3, channel, **ca), ResidualBlock(channel, channel, **ca))
        self.pool5_7x7 = nn.AvgPool2d(kernel_size=[7, 7], stride=[1, 1], padding=0)
        self.classifier = AddMarginProduct(channel)
        self._packed = None
        self._ws = {}
        self._profile = None
        # eval forward: spatial branch on a side stream (see _forward_eval). Opt-in (FFR_RECNET_BRANCH_STREAMS=1): measured
        # at batch 512 it changes nothing (9.528 vs 9.528 ms per step, tools/ab_bench.py 512 branch) - every kernel is a
        # persistent grid that fills the SMs, so the branches time-slice instead of overlapping
        self.branch_streams = os.environ.get("FFR_RECNET_BRANCH_STREAMS", "0") == "1"
        self.feat_space_mma = os.environ.get("FFR_FEAT_SPACE_MMA", "1") != "0"      # 0: the fp32 SIMT kernel

    # ------------------------------------------------------------------------------------------
    def conv_layers(self):
        """(name, ConvLayer) in execution order."""
        out = []
        for seq_name in ("Conv4Space", "ChannelFlipMerge", "Conv4Merge"):
            for i, m in enumerate(getattr(self, seq_name)):
                if isinstance(m, ConvLayer):
                    out.append(("%s.%d" % (seq_name, i), m))
                elif isinstance(m, ResidualBlock):
                    out.append(("%s.%d.conv1" % (seq_name, i), m.conv1))
                    out.append(("%s.%d.conv2" % (seq_name, i), m.conv2))
        return out

    def _cache_key(self):
        return (_lib.weights_generation(),) + tuple((t.data_ptr(), t._version)
                                                    for t in list(self.parameters()) + list(self.buffers()))

    def _pack_eval(self, device):
        key = (str(device),) + self._cache_key()
        if self._packed is not None and self._packed.key == key:
            return self._packed
        pk = _Obj()
        pk.key = key
        pk.conv = {}
        with torch.no_grad():
            for name, layer in self.conv_layers():
                bn = layer.norm.norm
                scale, shift = packing.bn_scale_shift(bn.weight.detach(), bn.bias.detach(), bn.running_mean,
                                                      bn.running_var, bn.eps)
                w = layer.conv2d.weight.detach()
                co, ci = w.shape[0], w.shape[1]
                cop, cip = _ceil64(co), _ceil64(ci)
                wp = packing.pack_conv(w, out_scale=scale, cin_pad=cip)
                c = _Obj()
                c.cin_p, c.cout_p = cip, cop
                c.wp = torch.zeros(cop, wp.shape[1], dtype=torch.bfloat16, device=device)
                c.wp[:co] = wp
                c.bias = torch.zeros(cop, dtype=torch.float32, device=device)
                c.bias[:co] = shift
                c.slope = torch.zeros(cop, dtype=torch.float32, device=device)
                c.slope[:co] = layer.relu.func.weight.detach().float()
                pk.conv[name] = c
            L = self.Conv4Channel
            w0 = L[0].weight.detach().float()                                  # [32, 561]
            pk.w0aT = w0[:, :49].t().contiguous()
            pk.w0bT = w0[:, 49:].t().contiguous()
            pk.b0 = L[0].bias.detach().float().contiguous()
            pk.slope1 = L[1].func.weight.detach().float().contiguous()
            pk.A1 = (L[3].weight.detach().float() @ L[2].weight.detach().float()).contiguous()
            pk.c1 = (L[3].weight.detach().float() @ L[2].bias.detach().float() + L[3].bias.detach().float()).contiguous()
            pk.slope4 = L[4].func.weight.detach().float().contiguous()
            pk.A2 = (L[6].weight.detach().float() @ L[5].weight.detach().float()).contiguous()
            pk.c2 = (L[6].weight.detach().float() @ L[5].bias.detach().float() + L[6].bias.detach().float()).contiguous()
            pk.slope7 = L[7].func.weight.detach().float().contiguous()
            w8 = torch.zeros(512, 64, dtype=torch.bfloat16, device=device)
            w8[:, :32] = L[8].weight.detach().to(torch.bfloat16)
            pk.w8 = w8
            pk.b8 = L[8].bias.detach().float().contiguous()
            pk.t_h9 = {off: _h9_scatter(off, device) for off in (0, 512)}
            pk.t_flip = _flip_scatter(device)
        self._packed = pk
        return pk

    def _workspace(self, n, device, slot=0):
        key = (n, str(device))
        held = self._ws.get(slot)
        if held is not None and held[0] == key:
            return held[1]
        ws = _Obj()
        bf = dict(dtype=torch.bfloat16, device=device)

        def h9(c):
            return torch.zeros(n * 81, c, **bf)
        ws.s0 = h9(576)
        ws.b256 = [h9(256) for _ in range(3)]
        ws.b128 = [h9(128) for _ in range(3)]
        ws.b64 = [h9(64) for _ in range(2)]
        ws.mspace = torch.zeros(n * 81, 64, dtype=torch.float32, device=device)
        ws.h5 = torch.zeros(n * 512, 64, **bf)
        ws.mch = torch.empty(n * 512, 512, **bf)
        ws.xt = torch.zeros(n * 128, 512, **bf)
        ws.fm = h9(1024)
        ws.cm = h9(1536)
        ws.c512 = [h9(512) for _ in range(2)]
        ws.d512 = [h9(512) for _ in range(3)]
        self._ws[slot] = (key, ws)
        return ws

    # ------------------------------------------------------------------------------------------
    def forward(self, input, label=None):
        """input: (N,512,7,7) fp32 CUDA. label None -> (feat_new_v (N,512), feat_new (N,512,7,7))  [recnet.py:425-426]."""
        if not input.is_cuda:
            raise RuntimeError("ffr_net_b200.RecNet runs only on CUDA (sm_100a); there is no CPU fallback")
        if input.dim() != 4 or tuple(input.shape[1:]) != (512, 7, 7):
            raise ValueError("expected input (N,512,7,7), got %s" % (tuple(input.shape),))
        if self.training:
            from . import recnet_train
            return recnet_train.forward_train(self, input, label)
        if torch.is_grad_enabled() and input.requires_grad:
            raise NotImplementedError("RecNet in eval mode is forward-only; a gradient w.r.t. its input is not implemented")
        if input.shape[0] == 0:                      # empty batch: nothing to launch
            return (torch.empty(0, 512, dtype=torch.float32, device=input.device),
                    torch.empty(0, 512, 7, 7, dtype=torch.float32, device=input.device))
        if label is not None:
            return self._forward_eval_label(input, label)
        n = input.shape[0]
        x = input.contiguous().float()
        v = torch.empty(n, 512, dtype=torch.float32, device=x.device)
        feat_new = torch.empty(n, 512, 7, 7, dtype=torch.float32, device=x.device)
        streams.fork_join(streams.chunk_bounds(n), x.device,
                          lambda i, lo, hi: self._forward_eval(x[lo:hi], True, slot=i, out_v=v[lo:hi],
                                                               out_map=feat_new[lo:hi]))
        return v, feat_new

    def _forward_eval_label(self, input, label):
        """Eval-mode call WITH a label: the reference returns the 7-tuple in any mode (recnet.py:425-429). Folded-BN
        forward as without label; M_space / M_channel / feat_space / feat_channel are exported from the workspace, the
        two (N,10575) head outputs are the literal AddMarginProduct expressions."""
        from . import recnet_train
        lib = _lib.load()
        P, st = _lib.ptr, _lib.stream_ptr()
        x = input.contiguous().float()
        n, dev = x.shape[0], x.device
        aux = {}
        v, _ = self._forward_eval(x, want_map=False, slot=0, aux=aux)
        ws = self._workspace(n, dev, 0)
        m_space = torch.empty(n, 64, 7, 7, dtype=torch.float32, device=dev)
        _lib.check(lib.ffr_rows_to_nchw(P(ws.mspace), 1, 64, 0, None, None, P(m_space), n, 7, 9, 1, 81, 64, st), "M_space")
        m_space = m_space[:, :49].reshape(n, 49, 49).contiguous()
        m_channel = ws.mch.view(n, 512, 512).float()
        feat_channel = torch.empty(n, 512, 7, 7, dtype=torch.float32, device=dev)
        _lib.check(lib.ffr_rows_to_nchw(P(ws.cm), 0, 1536, 512, None, None, P(feat_channel), n, 7, 9, 1, 81, 512, st),
                   "feat_channel")
        with torch.no_grad():
            pred_loss, pred_label = recnet_train.add_margin_product(self.classifier.weight, v, label, self.classifier.s,
                                                                    self.classifier.m)
        return v, pred_loss, pred_label, m_space, m_channel, aux["feat_space"], feat_channel

    def embed_from_images(self, encoder, x):
        """encoder(x) -> RecNet -> rectified embedding, the per-image path of lfw_eval.calculate_distance:241-242.
        The batch is cut into chunks that run concurrently on side streams (streams.py): the tail wave of one chunk's
        kernel is filled by the other chunk's next kernel, and HBM-bound passes overlap tensor-bound ones."""
        n = x.shape[0]
        u8 = x.dtype == torch.uint8                # decoded HWC images (Backbone.forward_u8): channel swap, no flip
        x = x.contiguous() if u8 else x.contiguous().float()
        v = torch.empty(n, 512, dtype=torch.float32, device=x.device)
        if n == 0:
            return v

        def chunk(i, lo, hi):
            y, _, _ = encoder.forward_internal(x[lo:hi], want_y=True, slot=i)
            self._forward_eval(y, want_map=False, slot=i, out_v=v[lo:hi])
        streams.fork_join(streams.chunk_bounds(n), x.device, chunk)
        return v

    def _forward_eval(self, x, want_map=True, slot=0, out_v=None, out_map=None, aux=None):
        lib = _lib.load()
        P = _lib.ptr
        prof = self._profile

        def chk(rc, what):
            _lib.check(rc, what)
            if prof is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                prof.append((what, e))

        x = x.contiguous().float()
        n, dev = x.shape[0], x.device
        pk = self._pack_eval(dev)
        ws = self._workspace(n, dev, slot)
        st = _lib.stream_ptr()

        chk(lib.ffr_recnet_prep(P(x), n, P(pk.w0aT), P(pk.w0bT), P(pk.b0), P(pk.slope1), P(pk.A1), P(pk.c1),
                                P(pk.slope4), P(pk.A2), P(pk.c2), P(pk.slope7), P(ws.s0), P(ws.cm), P(ws.xt),
                                P(ws.h5), None, st), "recnet_prep")

        def conv(name, src, dst, res=None, sigmoid=False, scatter=None, out_f32=None, pool=None):
            c = pk.conv[name]
            tab = scatter if scatter is not None else pk.t_h9[0]
            chk(lib.ffr_recnet_convlayer_fwd(P(src), n, c.cin_p, P(c.wp), c.cout_p, P(c.bias), P(c.slope),
                                             P(res), (res.shape[1] if res is not None else 0), 1 if sigmoid else 0,
                                             P(dst), (dst.shape[1] if dst is not None else 0), P(tab), 4, 81,
                                             P(out_f32), P(pool), _lib.stream_ptr()), name)

        # The spatial branch (nine small convolutions, latency- / feed-bound, + feat_space) and the channel branch
        # (M_channel, feat_channel, three 512-wide convolutions) are independent until Conv4Merge (recnet.py:404-420) and
        # write disjoint channel slots of the Conv4Merge input; with `branch_streams` the spatial branch runs on a side
        # stream so that its CTAs fill the SMs the channel branch's tail waves leave idle (same kernels, same bits).
        def spatial_branch():
            # ---- spatial rectifier (recnet.py:362-371, 404-405) ----
            conv("Conv4Space.0", ws.s0, ws.b256[0])
            conv("Conv4Space.1.conv1", ws.b256[0], ws.b256[1])
            conv("Conv4Space.1.conv2", ws.b256[1], ws.b256[2], res=ws.b256[0])
            conv("Conv4Space.2", ws.b256[2], ws.b128[0])
            conv("Conv4Space.3.conv1", ws.b128[0], ws.b128[1])
            conv("Conv4Space.3.conv2", ws.b128[1], ws.b128[2], res=ws.b128[0])
            conv("Conv4Space.4", ws.b128[2], ws.b64[0])
            conv("Conv4Space.5.conv1", ws.b64[0], ws.b64[1])
            conv("Conv4Space.5.conv2", ws.b64[1], None, res=ws.b64[0], sigmoid=True, out_f32=ws.mspace)
            fs_nchw = None
            if aux is not None:
                fs_nchw = torch.empty(n, 512, 7, 7, dtype=torch.float32, device=dev)
                aux["feat_space"] = fs_nchw
            if self.feat_space_mma:           # warp-MMA kernel on the X^T matrix recnet_prep wrote (same stream, earlier)
                chk(lib.ffr_feat_space_xt(P(ws.xt), P(ws.mspace), P(ws.cm), P(fs_nchw), n, _lib.stream_ptr()), "feat_space")
            else:
                chk(lib.ffr_feat_space(P(x), P(ws.mspace), P(ws.cm), P(fs_nchw), n, _lib.stream_ptr()), "feat_space")

        def channel_branch():
            # ---- channel rectifier (recnet.py:372-386, 406, 410): M_channel = sigmoid(h5 W8^T + b8); M_channel @ X ----
            EPI = _lib.EPI
            d = _lib.ConvGemmDesc()                      # M_channel rows (sample, c) = sigmoid(h5 W8^T + b8), K = 64 (32 valid)
            d.a, d.a_rows, d.a_cols, d.a_ld = P(ws.h5), n * 512, 64, 64
            d.wp, d.Cin, d.Cout, d.ntaps = P(pk.w8), 64, 512, 1
            d.M, d.n_img = n * 512, n
            d.flags = EPI.BIAS | EPI.SIGMOID
            d.bias = P(pk.b8)
            d.out, d.ldo = P(ws.mch), 512
            d.num_splits = 1
            chk(lib.ffr_conv_gemm_ex(ctypes.byref(d), _lib.stream_ptr()), "M_channel")
            d = _lib.ConvGemmDesc()                      # feat_channel = M_channel @ X per sample: batched weight operand
            d.a, d.a_rows, d.a_cols, d.a_ld = P(ws.xt), n * 128, 512, 512
            d.wp, d.Cin, d.Cout, d.ntaps = P(ws.mch), 512, 512, 1
            d.M, d.rows_per_img, d.Wp, d.n_img = n * 128, 128, 128, n
            d.flags = EPI.GEOM | EPI.SCATTER             # rows (h*7+w) -> flip / cat slots + reflection mirrors (t_flip)
            d.out, d.ldo = P(ws.fm), 1024
            d.scatter, d.scatter_n, d.out_rows_per_img = P(pk.t_flip), 8, 81
            d.num_splits, d.b_rows_per_mtile, d.b_mtile_div = 1, 512, 1
            chk(lib.ffr_conv_gemm_ex(ctypes.byref(d), _lib.stream_ptr()), "feat_channel")

            # ---- flip merge (recnet.py:415-418) and final merge (:420-423) ----
            conv("ChannelFlipMerge.0", ws.fm, ws.c512[0])
            conv("ChannelFlipMerge.1.conv1", ws.c512[0], ws.c512[1])
            conv("ChannelFlipMerge.1.conv2", ws.c512[1], ws.cm, res=ws.c512[0], scatter=pk.t_h9[512])

        if self.branch_streams:
            streams.fork_join([(0, 0), (1, 1)], dev, lambda i, lo, hi: (channel_branch if i == 0 else spatial_branch)())
        else:
            spatial_branch()
            channel_branch()
        conv("Conv4Merge.0", ws.cm, ws.d512[0])
        conv("Conv4Merge.1.conv1", ws.d512[0], ws.d512[1])
        conv("Conv4Merge.1.conv2", ws.d512[1], ws.d512[2], res=ws.d512[0])

        # pool5_7x7 (:424) over the stored map in a fixed order (an epilogue reduction would need atomics: the pooled
        # embedding is bit-reproducible this way)
        v = out_v if out_v is not None else torch.empty(n, 512, dtype=torch.float32, device=dev)
        chk(lib.ffr_h9_avgpool_bf16(P(ws.d512[2]), ws.d512[2].shape[1], P(v), 512, n, 512, st), "avgpool")
        feat_new = None
        if want_map:
            feat_new = out_map if out_map is not None else torch.empty(n, 512, 7, 7, dtype=torch.float32, device=dev)
            chk(lib.ffr_rows_to_nchw(P(ws.d512[2]), 0, 512, 0, None, None, P(feat_new), n, 7, 9, 1, 81, 512, st),
                "export")
        return v, feat_new


# ----------------------------------------------------------------------------------------------------------
# selfSimilarity / cosine_sim (reference :220-236) — public helpers used by the trainer's loss.
# ----------------------------------------------------------------------------------------------------------
def cosine_sim(x1, x2, dim=1):
    from . import recnet_train
    return recnet_train.cosine_sim(x1, x2, dim)


def selfSimilarity(x):
    from . import recnet_train
    return recnet_train.self_similarity(x)
