"""IR-SE50 backbone — drop-in for `pretrain/model_ir_se50.py` (Backbone / ir_se_50_512) of the reference.

Same constructor arguments, forward signature and state_dict key layout (402 keys) as the reference
(/root/reference/pretrain/model_ir_se50.py:108-154); the forward pass runs entirely in the sm_100a library
(stem conv, tcgen05 implicit-GEMM convolutions with folded BatchNorm/PReLU epilogues, fused SE+residual, folded
head GEMM + L2 norm). The torch sub-modules below only hold parameters under the reference's names; their own
forward() is never used. The backbone is frozen/eval-only in the reference (models/trainer.py:62-63,75,79): calling
it in training mode, or with CPU tensors, raises — there is no fallback path.
"""
from collections import namedtuple

import torch
from torch import nn

from . import _lib, packing, streams


class Flatten(nn.Module):
    def forward(self, input):
        return input.view(input.size(0), -1)


def l2_norm(input, axis=1):
    return input / torch.norm(input, 2, axis, True)


class SEModule(nn.Module):
    """Parameter holder for the squeeze-excitation block (reference :18-36)."""

    def __init__(self, channels, reduction):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(channels, channels // reduction, kernel_size=1, padding=0, bias=False)
        self.relu = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(channels // reduction, channels, kernel_size=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()


class bottleneck_IR_SE(nn.Module):
    """Parameter holder for one IR-SE unit (reference :56-76)."""

    def __init__(self, in_channel, depth, stride):
        super().__init__()
        self.in_channel, self.depth, self.stride = in_channel, depth, stride
        if in_channel == depth:
            self.shortcut_layer = nn.MaxPool2d(1, stride)
        else:
            self.shortcut_layer = nn.Sequential(nn.Conv2d(in_channel, depth, (1, 1), stride, bias=False),
                                                nn.BatchNorm2d(depth))
        self.res_layer = nn.Sequential(
            nn.BatchNorm2d(in_channel),
            nn.Conv2d(in_channel, depth, (3, 3), (1, 1), 1, bias=False),
            nn.PReLU(depth),
            nn.Conv2d(depth, depth, (3, 3), stride, 1, bias=False),
            nn.BatchNorm2d(depth),
            SEModule(depth, 16))


class Bottleneck(namedtuple("Block", ["in_channel", "depth", "stride"])):
    """A named tuple describing a ResNet block."""


def get_block(in_channel, depth, num_units, stride=2):
    return [Bottleneck(in_channel, depth, stride)] + [Bottleneck(depth, depth, 1) for _ in range(num_units - 1)]


def get_blocks(num_layers):
    units = {50: (3, 4, 14, 3), 100: (3, 13, 30, 3), 152: (3, 8, 36, 3)}[num_layers]
    chans = ((64, 64), (64, 128), (128, 256), (256, 512))
    return [get_block(ci, co, n) for (ci, co), n in zip(chans, units)]


def _bn_fold(bn):
    return packing.bn_scale_shift(bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps)


class _Packed:
    """Device-resident folded/packed weights of one backbone + the cache key they were built from."""
    pass


class _CpuModules:
    pass


def _cpu_view(model):
    """The module tree with every parameter / buffer copied to the host (ONE device->host transfer per tensor; a model
    that still lives on the CPU is used as is)."""
    import copy
    if all(not t.is_cuda for t in list(model.parameters()) + list(model.buffers())):
        return model
    packed, ws, prof = model._packed, model._ws, model._profile
    model._packed, model._ws, model._profile = None, {}, None
    try:
        clone = copy.deepcopy(model).cpu()
    finally:
        model._packed, model._ws, model._profile = packed, ws, prof
    return clone


def _to_device(pk, device):
    for obj in [pk] + list(getattr(pk, "units", [])):
        for k, v in list(vars(obj).items()):
            if torch.is_tensor(v):
                setattr(obj, k, v.to(device).contiguous())


class Backbone(nn.Module):
    IMG = 112

    def __init__(self, num_layers, drop_ratio, mode="ir"):
        super().__init__()
        assert num_layers in [50, 100, 152], "num_layers should be 50,100, or 152"
        assert mode in ["ir", "ir_se"], "mode should be ir or ir_se"
        if mode != "ir_se":
            raise NotImplementedError("only the 'ir_se' units used by FFR-Net (models/trainer.py:59) are implemented")
        self.input_layer = nn.Sequential(nn.Conv2d(3, 64, (3, 3), 1, 1, bias=False), nn.BatchNorm2d(64), nn.PReLU(64))
        self.output_layer = nn.Sequential(nn.BatchNorm2d(512), nn.Dropout(drop_ratio), Flatten(),
                                          nn.Linear(512 * 7 * 7, 512), nn.BatchNorm1d(512))
        self.bn = nn.BatchNorm2d(512)
        modules = []
        for block in get_blocks(num_layers):
            for b in block:
                modules.append(bottleneck_IR_SE(b.in_channel, b.depth, b.stride))
        self.body = nn.Sequential(*modules)
        self._packed = None
        self._ws = {}
        self._profile = None      # set to a list to collect (label, start_event, end_event) per library call
        self.fuse_se = True       # SE gate computed inside the scale + residual kernel (C <= 256); False: two launches

    # ------------------------------------------------------------------------------------------
    def _cache_key(self):
        # The backbone is frozen (models/trainer.py:62-63): its packed weights depend only on its own tensors. (The
        # fused optimizer's generation counter is deliberately NOT part of the key: it only ever updates RecNet.)
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _pack(self, device):
        key = (str(device),) + self._cache_key()
        if self._packed is not None and self._packed.key == key:
            return self._packed
        pk = _Packed()
        pk.key = key
        # The backbone is frozen: folding / packing happens once per load, on the HOST (plain torch CPU ops on copies of
        # the parameters), and only the packed tensors are uploaded — no device kernels are spent on weight preparation.
        cpu = _cpu_view(self)
        with torch.no_grad():
            conv, bn, prelu = cpu.input_layer[0], cpu.input_layer[1], cpu.input_layer[2]
            pk.stem_w, pk.stem_b = packing.pack_stem(conv.weight.detach(), _bn_fold(bn))
            pk.stem_a = prelu.weight.detach().float().contiguous()
            pk.units = []
            for unit in cpu.body:
                u = _Packed()
                u.cin, u.depth, u.stride = unit.in_channel, unit.depth, unit.stride
                bn0, conv1, pr, conv2, bn1, se = unit.res_layer
                s0, b0 = _bn_fold(bn0)
                s1, b1 = _bn_fold(bn1)
                u.w1 = packing.pack_conv(conv1.weight.detach(), in_scale=s0)
                u.bias9 = packing.border_bias_table(conv1.weight.detach(), b0)
                u.slope = pr.weight.detach().float().contiguous()
                u.w2 = packing.pack_conv(conv2.weight.detach(), out_scale=s1)
                u.b2 = b1.contiguous()
                u.fc1 = se.fc1.weight.detach().float().reshape(u.depth // 16, u.depth).contiguous()
                u.fc2 = se.fc2.weight.detach().float().reshape(u.depth, u.depth // 16).contiguous()
                if u.cin != u.depth:
                    ss, sb = _bn_fold(unit.shortcut_layer[1])
                    u.wsc = packing.pack_conv(unit.shortcut_layer[0].weight.detach(), out_scale=ss)
                    u.bsc = sb.contiguous()
                pk.units.append(u)
            pk.bn_scale, pk.bn_shift = [t.contiguous() for t in _bn_fold(cpu.bn)]
            lin, bn1d = cpu.output_layer[3], cpu.output_layer[4]
            pk.head_w, pk.head_b = packing.pack_head(lin.weight.detach(), lin.bias.detach(),
                                                     _bn_fold(cpu.output_layer[0]), _bn_fold(bn1d))
        _to_device(pk, device)
        self._packed = pk
        return pk

    def _workspace(self, n, device, slot=0):
        key = (n, str(device))
        held = self._ws.get(slot)
        if held is not None and held[0] == key:
            return held[1]
        ws = _Packed()
        S = self.IMG
        big = n * (S + 1) * (S + 1) * 64                     # elements of the largest flat map (112x112x64)
        bf = dict(dtype=torch.bfloat16, device=device)
        ws.a = torch.empty(big, **bf)
        ws.b = torch.empty(big, **bf)
        ws.t = torch.empty(big, **bf)
        ws.u = torch.empty(big // 2, **bf)
        ws.xs = torch.empty(big // 4, **bf)
        ws.sc = torch.empty(big // 4, **bf)
        # space-to-depth inputs of the stride-2 convs: their pad rows are never written, so they are zeroed once
        ws.s2d = {}
        res = S
        for i, unit in enumerate(self.body):
            if unit.stride == 2:
                so = res // 2
                ws.s2d[i] = torch.zeros(n * (so + 1) * (so + 1) * 4 * unit.depth, **bf)
                res = so
        lib = _lib.load()
        res, part = S, 0
        for unit in self.body:                                # SE squeeze partial sums: the largest layer's need
            res //= unit.stride
            part = max(part, lib.ffr_se_pool_part_floats(n, res, unit.depth))
        ws.pool_part = torch.empty(part, dtype=torch.float32, device=device)
        ws.gate = torch.empty(n * 512, dtype=torch.float32, device=device)
        ws.acc = torch.empty(lib.ffr_head_workspace_floats(n, res, 512), dtype=torch.float32, device=device)
        # stream-K scratch of the 256-wide 3x3 convolutions (flag words first: must start out zero)
        ws.sk = torch.zeros(lib.ffr_conv_scratch_bytes(), dtype=torch.uint8, device=device)
        self._ws[slot] = (key, ws)                            # keep one batch size resident per stream slot
        return ws

    # ------------------------------------------------------------------------------------------
    def forward(self, x):
        """x: (N,3,112,112) fp32 CUDA -> (featmap y (N,512,7,7) fp32, feat f (N,512) fp32, L2-normalised)."""
        if self.training:
            raise RuntimeError("Backbone is frozen/forward-only (reference: models/trainer.py:75,79); call .eval()")
        if not x.is_cuda:
            raise RuntimeError("ffr_net_b200.Backbone runs only on CUDA (sm_100a); there is no CPU fallback")
        if x.dtype == torch.uint8:
            return self.forward_u8(x)
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != self.IMG or x.shape[3] != self.IMG:
            raise ValueError("expected input (N,3,112,112), got %s" % (tuple(x.shape),))
        if x.shape[0] == 0:                          # empty batch: nothing to launch
            return (torch.empty(0, 512, 7, 7, dtype=torch.float32, device=x.device),
                    torch.empty(0, 512, dtype=torch.float32, device=x.device))
        n = x.shape[0]
        y = torch.empty(n, 512, 7, 7, dtype=torch.float32, device=x.device)
        f = torch.empty(n, 512, dtype=torch.float32, device=x.device)
        x = x.contiguous().float()
        streams.fork_join(streams.chunk_bounds(n), x.device,
                          lambda i, lo, hi: self.forward_internal(x[lo:hi], slot=i, out_y=y[lo:hi], out_f=f[lo:hi]))
        return y, f

    def forward_u8(self, img, flip=None, swap_rb=True):
        """Decoded images in, preprocessing fused into the stem: img uint8 (N,112,112,3) HWC CUDA as PIL decodes them
        (RGB); the reference's host pipeline — R/B channel swap (data/dataset.py:138-141), optional per-image
        horizontal flip (dataset.py:149-152; `flip`: uint8/bool (N,) CUDA or None), ToTensor + Normalize(0.5, 0.5)
        (data/dataloader.py:15-19) — runs inside the first kernel. Same outputs as forward(preprocessed fp32 NCHW)."""
        if self.training:
            raise RuntimeError("Backbone is frozen/forward-only (reference: models/trainer.py:75,79); call .eval()")
        if not img.is_cuda or img.dtype != torch.uint8:
            raise RuntimeError("forward_u8 expects a CUDA uint8 tensor; there is no CPU fallback")
        if img.dim() != 4 or tuple(img.shape[1:]) != (self.IMG, self.IMG, 3):
            raise ValueError("expected uint8 input (N,112,112,3), got %s" % (tuple(img.shape),))
        n = img.shape[0]
        y = torch.empty(n, 512, 7, 7, dtype=torch.float32, device=img.device)
        f = torch.empty(n, 512, dtype=torch.float32, device=img.device)
        if n == 0:
            return y, f
        img = img.contiguous()
        if flip is not None:
            flip = flip.to(device=img.device, dtype=torch.uint8).contiguous()
        streams.fork_join(streams.chunk_bounds(n), img.device,
                          lambda i, lo, hi: self.forward_internal(img[lo:hi], slot=i, out_y=y[lo:hi], out_f=f[lo:hi],
                                                                  flip=None if flip is None else flip[lo:hi],
                                                                  swap_rb=swap_rb))
        return y, f

    def forward_internal(self, x, want_y=True, slot=0, out_y=None, out_f=None, flip=None, swap_rb=True):
        """One chunk of images on the current stream. Returns (y, f, h) where h is the flat bf16 body output
        (N*64 rows x 512). `slot` picks the workspace (one per concurrent stream, streams.py); `out_y` / `out_f` are
        optional preallocated (contiguous) destinations."""
        lib = _lib.load()
        P = _lib.ptr
        prof = self._profile

        class L:                                  # per-call error check (+ optional CUDA-event timing)
            @staticmethod
            def check(rc, what=""):
                _lib.check(rc, what)
                if prof is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    prof.append((what, e))

        u8 = x.dtype == torch.uint8                # decoded HWC images: preprocessing fused into the stem
        x = x.contiguous() if u8 else x.contiguous().float()
        n, dev = x.shape[0], x.device
        pk = self._pack(dev)
        ws = self._workspace(n, dev, slot)
        st = _lib.stream_ptr()
        S = self.IMG
        if u8:
            L.check(lib.ffr_stem_u8_fwd(P(x), P(flip), 1 if swap_rb else 0, P(pk.stem_w), P(pk.stem_b), P(pk.stem_a),
                                        P(ws.a), n, S, st), "stem")
        else:
            L.check(lib.ffr_stem_fwd(P(x), P(pk.stem_w), P(pk.stem_b), P(pk.stem_a), P(ws.a), n, S, st), "stem")
        cur, nxt = ws.a, ws.b
        # stream-K scratch of this workspace slot for the convolution launches below (unregistered afterwards: the library
        # keeps only the pointer)
        _lib.check(lib.ffr_set_conv_scratch(P(ws.sk), ws.sk.numel()), "set_conv_scratch")
        try:
            for i, u in enumerate(pk.units):
                so = S // u.stride
                if u.stride == 2:
                    t = ws.s2d[i]
                    L.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(cur), n, S, u.cin, P(u.w1), u.depth, P(u.bias9), P(u.slope),
                                                            P(t), 1, st), "conv1 %d>%d@%ds1" % (u.cin, u.depth, S))
                else:
                    t = ws.t
                    L.check(lib.ffr_conv3x3_bnpre_prelu_fwd(P(cur), n, S, u.cin, P(u.w1), u.depth, P(u.bias9), P(u.slope),
                                                            P(t), 0, st), "conv1 %d>%d@%ds1" % (u.cin, u.depth, S))
                L.check(lib.ffr_conv3x3_bn_pool_fwd(P(t), n, S, u.depth, u.stride, P(u.w2), u.depth, P(u.b2), P(ws.u),
                                                    P(ws.pool_part), st), "conv2 %d>%d@%ds%d" % (u.depth, u.depth, S, u.stride))
                fused_se = self.fuse_se and u.depth <= 256      # gate + scale + residual in one launch (one CTA per image)
                if not fused_se:
                    L.check(lib.ffr_se_gate_fwd(P(ws.pool_part), P(u.fc1), P(u.fc2), P(ws.gate), None, n, so, u.depth, st), "se_gate")
                if u.cin == u.depth:
                    sc, mode = cur, (1 if u.stride == 2 else 0)
                else:
                    L.check(lib.ffr_subsample2(P(cur), P(ws.xs), n, so, u.cin, st), "subsample")
                    L.check(lib.ffr_conv1x1_bn_fwd(P(ws.xs), n, so, u.cin, P(u.wsc), u.depth, P(u.bsc), P(ws.sc), st),
                            "shortcut")
                    sc, mode = ws.sc, 2
                if fused_se:
                    L.check(lib.ffr_se_gate_residual_fwd(P(ws.u), P(ws.pool_part), P(u.fc1), P(u.fc2), P(sc), mode, P(nxt), n, so,
                                                         u.depth, st), "se_gate_residual")
                else:
                    L.check(lib.ffr_se_residual_fwd(P(ws.u), P(ws.gate), P(sc), mode, P(nxt), n, so, u.depth, st), "se_residual")
                cur, nxt = nxt, cur
                S = so
        finally:
            lib.ffr_set_conv_scratch(None, 0)
        y = None
        if want_y:
            y = out_y if out_y is not None else torch.empty(n, 512, S, S, dtype=torch.float32, device=dev)
            L.check(lib.ffr_export_nchw_fwd(P(cur), P(pk.bn_scale), P(pk.bn_shift), P(y), n, S, 512, st), "export")
        f = out_f if out_f is not None else torch.empty(n, 512, dtype=torch.float32, device=dev)
        L.check(lib.ffr_head_fwd(P(cur), n, S, 512, P(pk.head_w), P(pk.head_b), P(ws.acc), P(f), st), "head")
        return y, f, cur


def ir_se_50_512(weights_path="./pretrain/se50.pth", **kwargs):
    """Reference :143-154. Builds IR-SE50 and strictly loads `weights_path` if given."""
    model = Backbone(num_layers=50, drop_ratio=0.6, mode="ir_se")
    if weights_path:
        model.load_state_dict(torch.load(weights_path))
    return model
