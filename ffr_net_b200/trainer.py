"""Training step — mirror of the hot part of `models/trainer.py` (TripletLoss :31-43, Trainer.forward :139-152,
backward :154-180, optimizer_parameters :182-187, update_learning_rate :235-238) for one process per GPU.

Data parallelism replaces the reference's single-process `nn.parallel.data_parallel` (trainer.py:70,72): every rank
runs the frozen encoder and RecNet on its batch shard, the 76 RecNet/head gradient tensors are averaged with ONE NCCL
all-reduce over a flat bucket before clip_grad_value_ and Adam (losses are batch means, so the mean of per-rank
gradients equals the gradient of the global mean; BatchNorm statistics stay per rank like per-replica DP).
"""
import types

import torch
import torch.nn.functional as F
from torch import nn, optim

from . import _lib
from .backbone import Backbone
from .head import FusedCE
from .optim import FusedClipAdam
from .recnet import RecNet, init_weights, selfSimilarity


class TripletLoss(nn.Module):
    def forward(self, x_feat, y_feat, z_feat):
        margin = 0.1
        pos_cos = 1 - torch.sum(F.normalize(x_feat) * F.normalize(y_feat), 1)
        neg_cos = 1 - torch.sum(F.normalize(x_feat) * F.normalize(z_feat), 1)
        return F.relu((pos_cos - neg_cos) + margin).mean(), pos_cos.mean(), neg_cos.mean()


def default_opts(**kw):
    """The hyper-parameters run.py passes (run.py:10-28)."""
    o = types.SimpleNamespace(phase="train", lr=0.1, beta1=0.9, beta2=0.999, weight_decay=0.0, optimizer="adam",
                              loss_weight=[1.0, 1.0, 1.0, 1.0], device="cuda", continue_train=False,
                              ckpt_dir="./checkpoints", fused_head=True, two_streams=True,
                              merge_encoder_batches=True)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Trainer:
    def __init__(self, opts, encoder=None, recnet=None, encoder_weights=None):
        self.opts = opts
        self.isTrain = opts.phase.lower() == "train"
        self.lr = opts.lr
        dev = torch.device(opts.device)
        self.encoder = encoder if encoder is not None else Backbone(50, 0.6, "ir_se")
        if encoder_weights is not None:
            self.encoder.load_state_dict(encoder_weights)
        if recnet is None:
            recnet = RecNet(norm_type="bn", relu_type="prelu")
            init_weights(recnet, "kaiming")
        self.recnet = recnet
        for p in self.encoder.parameters():
            p.requires_grad = False
        self.encoder.to(dev).eval()
        self.recnet.to(dev)
        self._flat, self._flat_bound = None, False
        if self.isTrain:
            self.recnet.train()
            params = [p for p in self.recnet.parameters() if p.requires_grad]
            if opts.optimizer.lower() != "adam":
                raise NotImplementedError("run.py uses Adam (run.py:11)")
            # clip_grad_value_(1.0) + Adam fused into one kernel launch over all RecNet/head tensors
            self.optim = FusedClipAdam(params, opts.lr, betas=(opts.beta1, opts.beta2),
                                       weight_decay=opts.weight_decay, clip_value=1.0)
            self.sch = optim.lr_scheduler.MultiStepLR(self.optim, [5000, 10000, 15000], gamma=0.5)
            if dev.type == "cuda":
                self.bind_flat_gradients()         # every .grad is a view of one flat buffer from the start
        else:
            self.recnet.eval()
        # The thin library-op remainder of the step (Conv4Channel MLP, per-sample matmuls, CosFace head; see
        # recnet_train.py) would otherwise run as fp32 SIMT GEMMs: let cuBLAS use TF32 tensor cores for it.
        if getattr(opts, "tf32_glue", True):
            torch.backends.cuda.matmul.allow_tf32 = True
        # CosFace head + CrossEntropy as one fused forward/backward (head.py); opts.fused_head=False keeps the
        # reference's op sequence on library kernels (two (N,10575) tensors + nn.CrossEntropyLoss)
        self.fused_head = bool(getattr(opts, "fused_head", True))
        # run the two RecNet calls of an iteration on two streams (see _forward_two_streams)
        self.two_streams = bool(getattr(opts, "two_streams", True))
        self.merge_encoder_batches = bool(getattr(opts, "merge_encoder_batches", True))
        self._side = None
        self.mse_loss = nn.MSELoss()
        self.triplet = TripletLoss()
        self.cross_entropy = nn.CrossEntropyLoss()
        self._graph = None
        self._graph_opt = None
        self._static = None

    # ------------------------------------------------------------------------------------------------------
    # One whole training iteration. With capture_step() the iteration (2 encoder forwards, 2 RecNet forwards,
    # losses, backward, fused clip+Adam: ~1500 kernel launches) is recorded once into a CUDA graph and replayed,
    # which removes the host launch overhead; the learning rate and Adam step count live on the device.
    # ------------------------------------------------------------------------------------------------------
    def step(self, img1, img2, label):
        if self._graph is None:
            self.set_input(img1, img2, label)
            self.forward()
            self.optimizer_parameters(0)
        else:
            for dst, src in zip(self._static, (img1, img2, label)):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
            if self._graph_opt is not None:            # data parallel: the exchange runs between the two graphs
                self.allreduce_gradients()
                self._graph_opt.replay()
        self.update_learning_rate()

    def zero_grad(self):
        """One memset of the flat gradient buffer (76 tensors are views of it); per-tensor zeroing before it is bound."""
        if self._flat_bound and self._grads_bound():
            self._flat.zero_()
        elif self._flat_bound:                         # someone replaced / dropped a .grad (e.g. zero_grad(set_to_none=True))
            self.bind_flat_gradients()
        else:
            self.optim.zero_grad(set_to_none=False)

    def _grads_bound(self):
        off, base, ok = 0, self._flat.data_ptr(), True
        for p in self.recnet.parameters():
            if not p.requires_grad:
                continue
            ok = ok and p.grad is not None and p.grad.data_ptr() == base + 4 * off
            off += p.numel()
        return ok

    def bind_flat_gradients(self):
        """Make every RecNet/head .grad a view into ONE flat fp32 buffer: the data-parallel exchange is then a single
        all-reduce of that buffer with no pack / unpack copies (autograd accumulates into existing .grad in place)."""
        params = [p for p in self.recnet.parameters() if p.requires_grad]
        n = sum(p.numel() for p in params)
        self._flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self._flat_bound = True

    def capture_step(self, img1, img2, label, warmup=3, split_optimizer=None):
        """Record the iteration into CUDA graphs. One GPU: a single graph (forward, losses, backward, clip+Adam).
        Data parallel: graph 1 = forward + backward into the flat gradient buffer, then the NCCL all-reduce of that
        buffer runs eagerly on the same stream, then graph 2 = clip+Adam (models/trainer.py:182-187 order:
        backward -> [reduce] -> clip -> step)."""
        import torch.distributed as dist
        dp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if split_optimizer is not None:                # tests: exercise the two-graph path on one GPU
            dp = bool(split_optimizer)
        if dp and not self._flat_bound:
            self.bind_flat_gradients()
        self._static = (img1.clone(), img2.clone(), label.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.set_input(*self._static)
                self.forward()
                self.optimizer_parameters(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.set_input(*self._static)
            self.forward()
            if dp:
                self.zero_grad()
                self.backward()
            else:
                self.optimizer_parameters(0)
        if dp:
            graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph_opt, pool=graph.pool()):
                self.optim.step()
            self._graph_opt = graph_opt
        self._graph = graph

    def set_input(self, img1, img2, label):
        self.nonocl, self.ocl, self.gt_label = img1, img2, label

    def forward(self):
        with torch.no_grad():
            if self.merge_encoder_batches and self.nonocl.shape == self.ocl.shape:
                # the frozen eval-mode backbone is per-image: one forward over both image sets (better tile / wave
                # occupancy than two half-size ones), then split — same values as the two calls of trainer.py:141-142
                n = self.nonocl.shape[0]
                y, f = self.encoder(torch.cat((self.nonocl, self.ocl), 0))
                self.feat_map_non, self.feat_map_ocl = y[:n], y[n:]
                self.feat_extract_non, self.feat_extract_ocl = f[:n], f[n:]
            else:
                self.feat_map_non, self.feat_extract_non = self.encoder(self.nonocl)
                self.feat_map_ocl, self.feat_extract_ocl = self.encoder(self.ocl)
        if self.fused_head and self.recnet.training:
            from .recnet_train import forward_train
            rec = lambda fmap: forward_train(self.recnet, fmap, self.gt_label, fused_ce=True)
        else:
            rec = lambda fmap: self.recnet(fmap, self.gt_label)
        if self.two_streams and self.recnet.training:
            out_non, out_ocl = self._forward_two_streams(rec)
        else:
            out_non, out_ocl = rec(self.feat_map_non), rec(self.feat_map_ocl)
        (self.f_non, self.pred_loss_non, self.pred_label_non, self.M_space_non, self.M_channel_non, self.space_non,
         self.channel_non) = out_non
        (self.f_ocl, self.pred_loss_ocl, self.pred_label_ocl, self.M_space_ocl, self.M_channel_ocl, self.space_ocl,
         self.channel_ocl) = out_ocl
        if isinstance(self.pred_label_ocl, FusedCE):
            pred = self.pred_label_ocl.pred
        else:
            pred = self.pred_label_ocl.detach().argmax(1)
        self.pred_label = pred
        self._correct = pred.eq(self.gt_label).sum()          # no host sync here; .item() in get_current_values

    def _forward_two_streams(self, rec):
        """The two RecNet calls of an iteration (unmasked, masked; models/trainer.py:144-145) are independent until the
        losses: the second runs on a side stream forked from the current one, so its many small kernels (and, through
        autograd, their backward counterparts) overlap the first call's. Shared state is kept race-free: weights are
        packed before the fork, BatchNorm running statistics are applied after the join in call order."""
        from . import recnet_train
        dev = self.feat_map_non.device
        recnet_train.prepack(self.recnet)
        main = torch.cuda.current_stream(dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        fork = torch.cuda.Event()
        fork.record(main)
        self._side.wait_event(fork)
        with recnet_train.deferred_running_stats() as stats:
            with torch.cuda.stream(self._side):
                out_ocl = rec(self.feat_map_ocl)
                join = torch.cuda.Event()
                join.record(self._side)
            n_side = len(stats.items)
            out_non = rec(self.feat_map_non)
        main.wait_event(join)
        # reference order: all layers of the unmasked call, then all layers of the masked call
        stats.items = stats.items[n_side:] + stats.items[:n_side]
        stats.apply()
        return out_non, out_ocl

    def backward(self):
        # trainer.py:157-161 calls selfSimilarity five times and discards one of the two Grams in four of them;
        # only the Gram that enters a loss term is computed here (same values, half the work)
        from .recnet_train import self_similarity_channel, self_similarity_space
        ss_space, ss_channel = selfSimilarity(self.feat_map_non)
        ss_space_non = self_similarity_space(self.space_non)
        ss_space_ocl = self_similarity_space(self.space_ocl)
        ss_channel_non = self_similarity_channel(self.channel_non)
        ss_channel_ocl = self_similarity_channel(self.channel_ocl)
        mse = self.mse_loss
        l_space = (mse(ss_space, ss_space_non) + mse(ss_space, ss_space_ocl)) / 2
        l_channel = (mse(ss_channel, ss_channel_non) + mse(ss_channel, ss_channel_ocl)) / 2
        items = [(l_space + l_channel) / 2]
        t, self.pos_loss, self.neg_loss = self.triplet(self.f_ocl, self.feat_extract_non, self.feat_extract_ocl)
        items.append(t)
        items.append((mse(self.f_non, self.feat_extract_non) + mse(self.f_ocl, self.feat_extract_non)) / 2)
        def ce(logits):
            return logits.loss if isinstance(logits, FusedCE) else self.cross_entropy(logits, self.gt_label)
        items.append(ce(self.pred_loss_non) / (1e-8 + self.opts.loss_weight[3]) + ce(self.pred_loss_ocl))
        self.loss_items = [l * w for l, w in zip(items, self.opts.loss_weight)]
        sum(self.loss_items).backward()

    def save_model(self, file_name, extra_info=None):
        """models/trainer.py:216-224 (`<ckpt_dir>/<file_name>.pth.gzip`)."""
        from . import checkpoint
        return checkpoint.save_model(self.recnet, self.optim, self.opts.ckpt_dir, file_name, extra_info)

    def load_model(self, file_name):
        """models/trainer.py:201-214; sets self.start_point like the reference."""
        from . import checkpoint
        self.start_point = checkpoint.load_model(self.recnet, self.opts.ckpt_dir, file_name,
                                                 map_location=self.opts.device)
        _lib.bump_weights_generation()

    def allreduce_gradients(self):
        """Average the RecNet/head gradients over ranks with one all-reduce of a flat fp32 bucket (in place when the
        gradients are views of it, bind_flat_gradients; otherwise packed into and unpacked from it)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        if getattr(self, "_flat_bound", False) and self._grads_bound():
            dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
            self._flat.div_(dist.get_world_size())
            return
        params = [p for p in self.recnet.parameters() if p.grad is not None]
        n = sum(p.numel() for p in params)
        if self._flat is None or self._flat.numel() != n:
            self._flat = torch.empty(n, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            self._flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        dist.all_reduce(self._flat, op=dist.ReduceOp.SUM)
        self._flat.div_(dist.get_world_size())
        off = 0
        for p in params:
            p.grad.copy_(self._flat[off:off + p.numel()].view_as(p.grad))
            off += p.numel()

    def optimizer_parameters(self, cur_iters=0):
        self.zero_grad()                               # keeps gradient storage (addresses are cached by the optimizer)
        self.backward()
        self.allreduce_gradients()                     # before clipping, as reduce-then-clip in the reference
        self.optim.step()                              # clip_grad_value_(1.0) + Adam, one fused launch

    def get_current_values(self):
        keys = ["SelfSimilarityLoss", "TripletLoss", "IdentityLoss", "ClassifierLoss"]
        d = {k: "{:.4f}".format(v.item()) for k, v in zip(keys, self.loss_items)}
        self.accuracy = self._correct.item() / self.pred_label.shape[0]
        d["TrainAcc"] = "{:.4f}".format(self.accuracy)
        return d

    def get_lr(self):
        """models/trainer.py:226-227."""
        return {"LR": "{:.6f}".format(self.lr)}

    def get_pos(self):
        """models/trainer.py:229-230: mean (1 - cos) of the positive pairs of the triplet term."""
        return self.pos_loss

    def get_neg(self):
        """models/trainer.py:232-233."""
        return self.neg_loss

    def update_learning_rate(self):
        self.sch.step()
        for g in self.optim.param_groups:
            self.lr = g["lr"]
        self.optim.sync_lr()
