"""Oracle (test infrastructure): fp32 CPU restatement of the FFR-Net training step — Trainer.forward / backward /
optimizer_parameters of /root/reference/models/trainer.py:31-43,139-187 — on top of oracle.backbone / oracle.recnet."""
import torch
import torch.nn.functional as F

from . import backbone as ob
from . import recnet as orr


def triplet_loss(x, y, z, margin=0.1):
    """TripletLoss.forward, trainer.py:38-43."""
    pos = 1 - torch.sum(F.normalize(x) * F.normalize(y), 1)
    neg = 1 - torch.sum(F.normalize(x) * F.normalize(z), 1)
    return F.relu((pos - neg) + margin).mean(), pos.mean(), neg.mean()


def losses_from_outputs(feat_map_non, e_non, e_ocl, out_non, out_ocl, label, loss_weight, selfsim):
    """Trainer.backward, trainer.py:154-178 (everything up to loss.backward()). out_* are RecNet's 7-tuples."""
    f_non, pl_non, _, _, _, space_non, channel_non = out_non
    f_ocl, pl_ocl, _, _, _, space_ocl, channel_ocl = out_ocl
    ss_space, ss_channel = selfsim(feat_map_non)
    ss_space_non, _ = selfsim(space_non)
    ss_space_ocl, _ = selfsim(space_ocl)
    _, ss_channel_non = selfsim(channel_non)
    _, ss_channel_ocl = selfsim(channel_ocl)
    mse = F.mse_loss
    l_space = (mse(ss_space, ss_space_non) + mse(ss_space, ss_space_ocl)) / 2
    l_channel = (mse(ss_channel, ss_channel_non) + mse(ss_channel, ss_channel_ocl)) / 2
    items = [(l_space + l_channel) / 2]
    items.append(triplet_loss(f_ocl, e_non, e_ocl)[0])
    items.append((mse(f_non, e_non) + mse(f_ocl, e_non)) / 2)
    items.append(F.cross_entropy(pl_non, label) / (1e-8 + loss_weight[3]) + F.cross_entropy(pl_ocl, label))
    items = [l * w for l, w in zip(items, loss_weight)]
    return items, sum(items)


def train_step(bsd, rsd, img1, img2, label, loss_weight=(1.0, 1.0, 1.0, 1.0), emulate_bf16=False):
    """One Trainer.forward + backward on CPU fp32. Returns (loss_items, grads{name: tensor}, new BN stats, accuracy).
    emulate_bf16 rounds the tensors the device path stores in bf16 (see oracle.recnet._Ctx)."""
    params = {k: v.clone().requires_grad_(True) for k, v in rsd.items() if v.is_floating_point() and
              not k.endswith("running_mean") and not k.endswith("running_var")}
    sd = dict(rsd)
    sd.update(params)
    with torch.no_grad():
        y_non, e_non = ob.backbone_forward(bsd, img1)
        y_ocl, e_ocl = ob.backbone_forward(bsd, img2)
    out_non, st1 = orr.recnet_forward(sd, y_non, label, training=True, return_stats=True, emulate_bf16=emulate_bf16)
    sd2 = dict(sd)
    sd2.update({k: v.detach() for k, v in st1.items()})
    out_ocl, st2 = orr.recnet_forward(sd2, y_ocl, label, training=True, return_stats=True, emulate_bf16=emulate_bf16)
    items, loss = losses_from_outputs(y_non, e_non, e_ocl, out_non, out_ocl, label, loss_weight, orr.self_similarity)
    loss.backward()
    acc = (out_ocl[2].argmax(1) == label).float().mean().item()
    grads = {k: p.grad for k, p in params.items()}
    return [float(i.detach()) for i in items], grads, {k: v.detach() for k, v in st2.items()}, acc


def clip_adam_step(params, grads, lr, step=1, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip=1.0, state=None):
    """clip_grad_value_(params, 1.0) followed by one torch.optim.Adam step (models/trainer.py:183-187), restated with
    the textbook update on plain tensors. `state` carries (exp_avg, exp_avg_sq) per name between steps."""
    state = {} if state is None else state
    out = {}
    b1, b2 = betas
    for k, p in params.items():
        g = grads[k].clamp(-clip, clip)
        if weight_decay != 0.0:
            g = g + weight_decay * p
        m, v = state.get(k, (torch.zeros_like(p), torch.zeros_like(p)))
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        state[k] = (m, v)
        denom = (v.sqrt() / (1 - b2 ** step) ** 0.5) + eps
        out[k] = p - (lr / (1 - b1 ** step)) * m / denom
    return out, state
