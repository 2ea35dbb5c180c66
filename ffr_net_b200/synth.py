"""Deterministic synthetic weights and inputs (the reference's checkpoints and datasets are Google-Drive hosted and
unavailable offline; BASELINE.json asks for synthetic data of the right shape). Shared by bench.py, the tests, the
oracle and the developer tools, so that the product benchmark never has to import oracle/. Pure torch / numpy on the
CPU; the golden fixtures of tests/golden/ were generated from exactly these generators (tools/make_golden.py)."""
import math

import torch

NUM_CLASSES = 10575


def unit_table(num_layers=50):
    """(in_channel, depth, stride) per bottleneck unit — get_block/get_blocks, model_ir_se50.py:81-91."""
    assert num_layers == 50
    out = []
    for cin, depth, n in ((64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)):
        out.append((cin, depth, 2))
        out += [(depth, depth, 1)] * (n - 1)
    return out


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


# ----------------------------------------------------------------------------------------------------------
# Deterministic synthetic weights: init_weights(recnet, 'kaiming') semantics (recnet.py:13-42, trainer.py:65-66):
# conv/linear weights kaiming-normal(fan_in), biases 0; BatchNorm weight ~ N(1, 0.02), bias 0; PReLU 0.25;
# classifier xavier-uniform. `perturb=True` additionally moves BN stats / PReLU slopes / linear biases away from
# their trivial values so that folding mistakes are visible.
# ----------------------------------------------------------------------------------------------------------
def synth_recnet_state_dict(seed=0, perturb=True):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def kaiming(shape, fan_in):
        return torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)

    for p, cin, cout in CONV_LAYERS:
        sd[p + ".conv2d.weight"] = kaiming((cout, cin, 3, 3), cin * 9)
        sd[p + ".relu.func.weight"] = (torch.empty(cout).uniform_(0.1, 0.4, generator=g) if perturb
                                       else torch.full((cout,), 0.25))
        q = p + ".norm.norm."
        sd[q + "weight"] = 1.0 + 0.02 * torch.randn(cout, generator=g)
        sd[q + "bias"] = torch.empty(cout).uniform_(-0.1, 0.1, generator=g) if perturb else torch.zeros(cout)
        sd[q + "running_mean"] = torch.empty(cout).uniform_(-0.1, 0.1, generator=g) if perturb else torch.zeros(cout)
        sd[q + "running_var"] = torch.empty(cout).uniform_(0.8, 1.2, generator=g) if perturb else torch.ones(cout)
        sd[q + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    for p, cin, cout in LINEARS:
        sd[p + ".weight"] = kaiming((cout, cin), cin)
        sd[p + ".bias"] = torch.empty(cout).uniform_(-0.1, 0.1, generator=g) if perturb else torch.zeros(cout)
    for i in (1, 4, 7):
        sd["Conv4Channel.%d.func.weight" % i] = (torch.empty(512).uniform_(0.1, 0.4, generator=g) if perturb
                                                 else torch.full((512,), 0.25))
    b = math.sqrt(6.0 / (512 + NUM_CLASSES))
    sd["classifier.weight"] = torch.empty(NUM_CLASSES, 512).uniform_(-b, b, generator=g)
    return sd
