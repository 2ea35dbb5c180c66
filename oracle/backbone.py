"""Oracle (test infrastructure): fp32 CPU restatement of the IR-SE50 backbone forward.

Follows /root/reference/pretrain/model_ir_se50.py — cited per function. Operates on a plain state_dict with the
reference's key layout so the same weights drive the reference, the oracle and the CUDA path.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d / BatchNorm1d default, model_ir_se50.py:64,66,70,119,121,125,126


def unit_table(num_layers=50):
    """(in_channel, depth, stride) per bottleneck unit — get_block/get_blocks, model_ir_se50.py:81-91."""
    assert num_layers == 50
    out = []
    for cin, depth, n in ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)):
        out.append((cin, depth, 2))
        out += [(depth, depth, 1)] * (n - 1)
    return out


def _bn(x, sd, p):
    """Eval-mode BatchNorm with running statistics."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, BN_EPS)


def se_module(x, sd, p):
    """SEModule.forward, model_ir_se50.py:29-36: x * sigmoid(fc2(relu(fc1(avgpool(x)))))."""
    m = x.mean(dim=(2, 3), keepdim=True)
    m = F.relu(F.conv2d(m, sd[p + "fc1.weight"]))
    m = torch.sigmoid(F.conv2d(m, sd[p + "fc2.weight"]))
    return x * m


def bottleneck_ir_se(x, sd, p, cin, depth, stride):
    """bottleneck_IR_SE.forward, model_ir_se50.py:56-76."""
    if cin == depth:
        shortcut = x[:, :, ::stride, ::stride]                       # MaxPool2d(1, stride), :59-60
    else:
        shortcut = _bn(F.conv2d(x, sd[p + "shortcut_layer.0.weight"], stride=stride), sd,
                       p + "shortcut_layer.1.")                       # :62-64
    r = _bn(x, sd, p + "res_layer.0.")                                # :66
    r = F.conv2d(r, sd[p + "res_layer.1.weight"], stride=1, padding=1)  # :67
    r = F.prelu(r, sd[p + "res_layer.2.weight"])                      # :68
    r = F.conv2d(r, sd[p + "res_layer.3.weight"], stride=stride, padding=1)  # :69
    r = _bn(r, sd, p + "res_layer.4.")                                # :70
    r = se_module(r, sd, p + "res_layer.5.")                          # :71
    return r + shortcut                                               # :76


def l2_norm(x, axis=1):
    """l2_norm, model_ir_se50.py:13-16 (no epsilon)."""
    return x / torch.norm(x, 2, axis, True)


def backbone_forward(sd, x, return_body=False):
    """Backbone.forward (eval), model_ir_se50.py:136-141. x: (N,3,112,112) fp32 -> (y (N,512,7,7), f (N,512))."""
    h = F.conv2d(x, sd["input_layer.0.weight"], padding=1)            # :118
    h = _bn(h, sd, "input_layer.1.")
    h = F.prelu(h, sd["input_layer.2.weight"])
    for u, (cin, depth, stride) in enumerate(unit_table()):
        h = bottleneck_ir_se(h, sd, "body.%d." % u, cin, depth, stride)
    y = _bn(h, sd, "bn.")                                             # :126,139
    o = _bn(h, sd, "output_layer.0.")                                 # :121  (Dropout is identity in eval, :122)
    o = o.reshape(o.size(0), -1)                                      # Flatten, :9-11
    o = F.linear(o, sd["output_layer.3.weight"], sd["output_layer.3.bias"])  # :124
    o = F.batch_norm(o, sd["output_layer.4.running_mean"], sd["output_layer.4.running_var"],
                     sd["output_layer.4.weight"], sd["output_layer.4.bias"], False, 0.0, BN_EPS)  # :125
    f = l2_norm(o)                                                    # :141
    if return_body:
        return y, f, h
    return y, f


# ----------------------------------------------------------------------------------------------------------
# Deterministic synthetic weights (checkpoints are Google-Drive hosted and unavailable offline).
# ----------------------------------------------------------------------------------------------------------
def _bn_entries(sd, p, c, g):
    sd[p + "weight"] = torch.empty(c).uniform_(0.8, 1.2, generator=g)
    sd[p + "bias"] = torch.empty(c).uniform_(-0.1, 0.1, generator=g)
    sd[p + "running_mean"] = torch.empty(c).uniform_(-0.1, 0.1, generator=g)
    sd[p + "running_var"] = torch.empty(c).uniform_(0.8, 1.2, generator=g)
    sd[p + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv_w(shape, g):
    fan_in = shape[1] * shape[2] * shape[3]
    b = 1.0 / math.sqrt(fan_in)          # == nn.Conv2d default kaiming_uniform_(a=sqrt(5)) bound
    return torch.empty(shape).uniform_(-b, b, generator=g)


def synth_backbone_state_dict(seed=0):
    """Random-init state_dict with the reference's 402 keys. BN affine/running stats and PReLU
    slopes are drawn away from identity so that folding bugs are visible."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    sd["input_layer.0.weight"] = _conv_w((64, 3, 3, 3), g)
    _bn_entries(sd, "input_layer.1.", 64, g)
    sd["input_layer.2.weight"] = torch.empty(64).uniform_(0.1, 0.4, generator=g)
    for u, (cin, depth, stride) in enumerate(unit_table()):
        p = "body.%d." % u
        if cin != depth:
            sd[p + "shortcut_layer.0.weight"] = _conv_w((depth, cin, 1, 1), g)
            _bn_entries(sd, p + "shortcut_layer.1.", depth, g)
        _bn_entries(sd, p + "res_layer.0.", cin, g)
        sd[p + "res_layer.1.weight"] = _conv_w((depth, cin, 3, 3), g)
        sd[p + "res_layer.2.weight"] = torch.empty(depth).uniform_(0.1, 0.4, generator=g)
        sd[p + "res_layer.3.weight"] = _conv_w((depth, depth, 3, 3), g)
        _bn_entries(sd, p + "res_layer.4.", depth, g)
        sd[p + "res_layer.5.fc1.weight"] = _conv_w((depth // 16, depth, 1, 1), g)
        sd[p + "res_layer.5.fc2.weight"] = _conv_w((depth, depth // 16, 1, 1), g)
    _bn_entries(sd, "output_layer.0.", 512, g)
    b = 1.0 / math.sqrt(25088)
    sd["output_layer.3.weight"] = torch.empty(512, 25088).uniform_(-b, b, generator=g)
    sd["output_layer.3.bias"] = torch.empty(512).uniform_(-b, b, generator=g)
    _bn_entries(sd, "output_layer.4.", 512, g)
    _bn_entries(sd, "bn.", 512, g)
    return sd


def synth_faces(n, seed=0, masked=False):
    """Synthetic 'face' batch in [-1,1] (range of ToTensor+Normalize(.5,.5), data/dataloader.py:15-19).
    masked=True overwrites rows 56..111 with a per-image, per-channel constant."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, 112, 112, generator=g).mul_(0.5).clamp_(-1, 1)
    if masked:
        g2 = torch.Generator().manual_seed(seed + 1)
        col = torch.empty(n, 3, 1, 1).uniform_(-1, 1, generator=g2)
        x[:, :, 56:, :] = col
    return x
