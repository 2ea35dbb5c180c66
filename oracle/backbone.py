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


# Deterministic synthetic weights / inputs live in ffr_net_b200/synth.py (shared by tests, bench and tools so that the
# product benchmark never imports oracle/); re-exported here under their historical names.
from ffr_net_b200.synth import synth_backbone_state_dict, synth_faces  # noqa: E402,F401
