"""Host-side weight preparation: BatchNorm folding and packing into the K-major bf16 layout the tcgen05 kernels read.

Done once per load_state_dict (backbone) or per optimizer step (RecNet); plain torch ops on whatever device the
parameters live on. Nothing here is on the per-batch path.
"""
import torch

BN_EPS = 1e-5


def bn_scale_shift(weight, bias, mean, var, eps=BN_EPS):
    """Eval BatchNorm as y = x*scale + shift."""
    scale = weight.float() / torch.sqrt(var.float() + eps)
    shift = bias.float() - mean.float() * scale
    return scale, shift


def pack_conv(w, in_scale=None, out_scale=None, cin_pad=None):
    """OIHW fp32 conv weight -> bf16 [Cout][kh*kw*Cin_p], k = (r*kw+s)*Cin_p + ci, optionally folding a per-input-
    channel scale (BatchNorm before the conv) and a per-output-channel scale (BatchNorm after it)."""
    w = w.float()
    if in_scale is not None:
        w = w * in_scale.view(1, -1, 1, 1)
    if out_scale is not None:
        w = w * out_scale.view(-1, 1, 1, 1)
    co, ci, kh, kw = w.shape
    w = w.permute(0, 2, 3, 1)  # O,H,W,I
    if cin_pad is not None and cin_pad != ci:
        wp = w.new_zeros(co, kh, kw, cin_pad)
        wp[..., :ci] = w
        w = wp
    return w.reshape(co, -1).to(torch.bfloat16).contiguous()


def border_bias_table(w, in_shift):
    """[9][Cout] fp32: contribution of a pre-conv BatchNorm shift under ZERO padding (padding is applied after the
    BatchNorm, model_ir_se50.py:66-67): for border class (ch,cw) only the taps that fall inside the image see the
    shift. Class index ch*3+cw with 0 = first row/col, 1 = interior, 2 = last row/col."""
    t = torch.einsum("oirs,i->ors", w.float(), in_shift.float())  # [Cout,3,3]
    rows = {0: (1, 2), 1: (0, 1, 2), 2: (0, 1)}
    out = []
    for ch in range(3):
        for cw in range(3):
            acc = 0
            for r in rows[ch]:
                for s in rows[cw]:
                    acc = acc + t[:, r, s]
            out.append(acc)
    return torch.stack(out, 0).contiguous()


def pack_head(lin_w, lin_b, bn2d, bn1d, S=7, C=512):
    """Folds output_layer (BatchNorm2d -> Flatten(NCHW) -> Linear -> BatchNorm1d, model_ir_se50.py:121-125) into a
    single bf16 GEMM over one image's flat rows: W' [512][(S+1)^2*C] (zero at pad pixels), b' [512] fp32."""
    so, bo = bn2d
    s1, b1 = bn1d
    W = lin_w.float().view(-1, C, S, S)                                   # [D, c, h, w]  (flatten index c*49+h*7+w)
    bias = s1 * (lin_b.float() + torch.einsum("dchw,c->d", W, bo)) + b1
    Wf = W * so.view(1, C, 1, 1) * s1.view(-1, 1, 1, 1)
    G = S + 1
    Wp = Wf.new_zeros(W.shape[0], G, G, C)
    Wp[:, :S, :S, :] = Wf.permute(0, 2, 3, 1)
    return Wp.reshape(W.shape[0], -1).to(torch.bfloat16).contiguous(), bias.contiguous()


def pack_stem(w, bn):
    """input_layer conv+BN (model_ir_se50.py:118-119) -> fp32 [27][64] (k = ci*9+r*3+s) and the shift [64]."""
    scale, shift = bn
    wf = w.float() * scale.view(-1, 1, 1, 1)          # [64,3,3,3]
    return wf.reshape(64, 27).t().contiguous(), shift.contiguous()
