"""Oracle (test infrastructure): fp32 CPU restatement of RecNet (feature rectification) forward.

Follows /root/reference/models/recnet.py — cited per function. Works on a plain state_dict with the reference's
121 keys. `training=True` uses batch statistics in every BatchNorm (and returns the updated
running statistics), as `recnet.train()` does in models/trainer.py:80.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
NUM_CLASSES = 10575

# ConvLayer prefixes in execution order with (Cin, Cout) — RecNet.__init__, recnet.py:356-396
CONV_LAYERS = [
    ("Conv4Space.0", 561, 256), ("Conv4Space.1.conv1", 256, 256), ("Conv4Space.1.conv2", 256, 256),
    ("Conv4Space.2", 256, 128), ("Conv4Space.3.conv1", 128, 128), ("Conv4Space.3.conv2", 128, 128),
    ("Conv4Space.4", 128, 49), ("Conv4Space.5.conv1", 49, 49), ("Conv4Space.5.conv2", 49, 49),
    ("ChannelFlipMerge.0", 1024, 512), ("ChannelFlipMerge.1.conv1", 512, 512), ("ChannelFlipMerge.1.conv2", 512, 512),
    ("Conv4Merge.0", 1536, 512), ("Conv4Merge.1.conv1", 512, 512), ("Conv4Merge.1.conv2", 512, 512),
]
LINEARS = [("Conv4Channel.0", 561, 32), ("Conv4Channel.2", 32, 512), ("Conv4Channel.3", 512, 32),
           ("Conv4Channel.5", 32, 512), ("Conv4Channel.6", 512, 32), ("Conv4Channel.8", 32, 512)]


def cosine_sim(x1, x2, dim=1):
    """cosine_sim, recnet.py:220-224 (F.normalize eps 1e-12 on dim 2, then bmm)."""
    x1 = F.normalize(x1, dim=2)
    x2 = F.normalize(x2, dim=2)
    return torch.bmm(x1, x2.permute(0, 2, 1))


def self_similarity(x):
    """selfSimilarity, recnet.py:226-236: (N,C,H,W) -> ss_space (N,HW,H,W), ss_channel (N,C,C)."""
    n, c, h, w = x.shape
    v = x.reshape(n, c, -1)
    ss_space = cosine_sim(v.permute(0, 2, 1), v.permute(0, 2, 1))
    ss_channel = cosine_sim(v, v)
    return ss_space.reshape(n, h * w, h, w), ss_channel


class _RoundBf16(torch.autograd.Function):
    """Identity that rounds to bf16 in forward AND backward: models a tensor (and its gradient) that the device
    path stores in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _Ctx:
    def __init__(self, sd, training, emulate_bf16=False):
        self.sd = sd
        self.training = training
        self.new_stats = {}
        # emulate_bf16: round ConvLayer inputs, weights and raw conv outputs (and the gradients flowing through the
        # same points) to bf16 exactly where the device path stores bf16 — separates "storage precision" from bugs
        self.emulate_bf16 = emulate_bf16


def _conv_layer(ctx, x, p):
    """ConvLayer.forward, recnet.py:78-85: ReflectionPad2d(1) -> Conv2d 3x3 (no bias) -> BatchNorm2d -> PReLU."""
    sd = ctx.sd
    w = sd[p + ".conv2d.weight"]
    if ctx.emulate_bf16:
        x = _RoundBf16.apply(x)
        w = w + (w.detach().bfloat16().float() - w.detach())        # bf16 weight values, fp32 gradient
    out = F.pad(x, (1, 1, 1, 1), mode="reflect")
    out = F.conv2d(out, w)
    if ctx.emulate_bf16:
        out = _RoundBf16.apply(out)
    q = p + ".norm.norm."
    if ctx.training:
        mean = out.mean(dim=(0, 2, 3))
        var = out.var(dim=(0, 2, 3), unbiased=False)
        cnt = out.numel() / out.shape[1]
        ctx.new_stats[q + "running_mean"] = (1 - BN_MOMENTUM) * sd[q + "running_mean"] + BN_MOMENTUM * mean
        ctx.new_stats[q + "running_var"] = (1 - BN_MOMENTUM) * sd[q + "running_var"] + \
            BN_MOMENTUM * var * cnt / (cnt - 1)
        ctx.new_stats[q + "num_batches_tracked"] = sd[q + "num_batches_tracked"] + 1
        out = (out - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
        out = out * sd[q + "weight"].view(1, -1, 1, 1) + sd[q + "bias"].view(1, -1, 1, 1)
    else:
        out = F.batch_norm(out, sd[q + "running_mean"], sd[q + "running_var"], sd[q + "weight"], sd[q + "bias"],
                           False, 0.0, BN_EPS)
    return F.prelu(out, sd[p + ".relu.func.weight"])


def _res_block(ctx, x, p):
    """ResidualBlock.forward, recnet.py:213-218 (the add comes after the second PReLU)."""
    return _conv_layer(ctx, _conv_layer(ctx, x, p + ".conv1"), p + ".conv2") + x


def _conv4space(ctx, x):
    """RecNet.Conv4Space, recnet.py:362-371."""
    h = _conv_layer(ctx, x, "Conv4Space.0")
    h = _res_block(ctx, h, "Conv4Space.1")
    h = _conv_layer(ctx, h, "Conv4Space.2")
    h = _res_block(ctx, h, "Conv4Space.3")
    h = _conv_layer(ctx, h, "Conv4Space.4")
    h = _res_block(ctx, h, "Conv4Space.5")
    return torch.sigmoid(h)


def _conv4channel(ctx, x):
    """RecNet.Conv4Channel, recnet.py:372-386. x: (N,512,561); nn.PReLU(512) acts on dim 1 (the 512 rows)."""
    sd = ctx.sd
    h = x
    for i in (0, 3, 6):
        h = F.linear(h, sd["Conv4Channel.%d.weight" % i], sd["Conv4Channel.%d.bias" % i])
        h = F.prelu(h, sd["Conv4Channel.%d.func.weight" % (i + 1)])
        h = F.linear(h, sd["Conv4Channel.%d.weight" % (i + 2)], sd["Conv4Channel.%d.bias" % (i + 2)])
    return torch.sigmoid(h)


def add_margin_product(sd, x, label, s=30.0, m=0.40):
    """AddMarginProduct.forward, recnet.py:257-270 (CosFace): returns (s*(cos - m*onehot), cos)."""
    cosine = F.linear(F.normalize(x), F.normalize(sd["classifier.weight"]))
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    output = (one_hot * (cosine - m)) + ((1.0 - one_hot) * cosine)
    return output * s, cosine


def recnet_forward(sd, x, label=None, training=False, return_stats=False, emulate_bf16=False):
    """RecNet.forward, recnet.py:398-429. x: (N,512,7,7) fp32."""
    ctx = _Ctx(sd, training, emulate_bf16)
    n, c, hh, ww = x.shape
    ss_space, ss_channel = self_similarity(x)                                   # :399
    space_cat = torch.cat((x, ss_space), 1)                                     # :401
    flat = x.reshape(n, c, -1)
    channel_cat = torch.cat((flat, ss_channel), 2)                              # :402
    m_space = _conv4space(ctx, space_cat).reshape(n, hh * ww, -1)               # :404-405
    m_channel = _conv4channel(ctx, channel_cat)                                 # :406
    feat_space = torch.matmul(flat, m_space).reshape(n, c, hh, ww)              # :409,412
    feat_channel = torch.matmul(m_channel, flat).reshape(n, c, hh, ww)          # :410,413
    fc_cat = torch.cat((torch.flip(feat_channel, [3]), feat_channel), 1)        # :416-417
    h = _conv_layer(ctx, fc_cat, "ChannelFlipMerge.0")                          # :418
    feat_channel = _res_block(ctx, h, "ChannelFlipMerge.1")
    feat_cat = torch.cat((feat_space, feat_channel, x), 1)                      # :420
    h = _conv_layer(ctx, feat_cat, "Conv4Merge.0")                              # :421
    feat_new = _res_block(ctx, h, "Conv4Merge.1")
    feat_new_v = feat_new.mean(dim=(2, 3))                                      # AvgPool2d(7), :423
    if label is None:
        out = (feat_new_v, feat_new)                                            # :426
    else:
        pred_loss, pred_label = add_margin_product(sd, feat_new_v, label)       # :428
        out = (feat_new_v, pred_loss, pred_label, m_space, m_channel, feat_space, feat_channel)
    if return_stats:
        return out, ctx.new_stats
    return out


# Deterministic synthetic weights live in ffr_net_b200/synth.py (see oracle/backbone.py); re-exported here.
from ffr_net_b200.synth import synth_recnet_state_dict  # noqa: E402,F401
