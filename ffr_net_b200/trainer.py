"""Training step — mirror of the hot part of `models/trainer.py` (TripletLoss :31-43, Trainer.__init__ :52-95,
clone_model :97-113, forward :139-152, backward :154-180, optimizer_parameters :182-187, update_learning_rate :235-238)
for one process per GPU.

The whole iteration is hand-written sm_100a code behind the C ABI: frozen IR-SE50 forward (backbone.py), RecNet forward /
backward for the unmasked and the masked batch as ONE batch of 2n samples (recnet_train.TrainEngine), the four losses
with their gradients (csrc/loss_kernels.cu, head.py), clip_grad_value_ + Adam in one launch (optim.py). No autograd graph
is built; the gradients land directly in one flat fp32 buffer that `param.grad` views alias.

Data parallelism replaces the reference's single-process `nn.parallel.data_parallel` (trainer.py:70,72): every rank
runs the frozen encoder and RecNet on its batch shard and the 76 RecNet/head gradient tensors are averaged over ranks
with NCCL before clip_grad_value_ and Adam (losses are batch means, so the mean of per-rank gradients equals the
gradient of the global mean; BatchNorm statistics stay per rank like per-replica DP). The flat buffer is exchanged in
buckets on a side stream as soon as the backward pass has produced them (classifier first, Conv4Space last), so the
all-reduce overlaps the rest of the backward pass.
"""
import ctypes
import types

import torch
import torch.nn.functional as F
from torch import nn, optim

from . import _lib, head, losses
from .backbone import Backbone
from .optim import FusedClipAdam
from .recnet import RecNet, init_weights, selfSimilarity
from . import recnet_train

EPI = _lib.EPI
_P = _lib.ptr


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
                              which_file="latest", ckpt_dir="./checkpoints", literal=False, merge_encoder_batches=True,
                              overlap_allreduce=True, data_parallel=True)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class _Obj:
    pass


# gradient buckets in the order the backward pass finishes them (prefixes of parameter names)
_BUCKET_ORDER = ("classifier", "Conv4Merge", "ChannelFlipMerge", "Conv4Channel", "Conv4Space")


class Trainer:
    def __init__(self, opts, encoder=None, recnet=None, encoder_weights=None):
        self.opts = opts
        self.isTrain = opts.phase.lower() == "train"
        self.norm_type, self.relu_type = "bn", "prelu"
        self.lr = opts.lr
        dev = torch.device(opts.device)
        self.encoder = encoder if encoder is not None else Backbone(50, 0.6, "ir_se")
        if encoder_weights is not None:
            self.encoder.load_state_dict(encoder_weights)
        if recnet is None:
            recnet = RecNet(norm_type=self.norm_type, relu_type=self.relu_type)
            init_weights(recnet, "kaiming")
        self.recnet = recnet
        for p in self.encoder.parameters():
            p.requires_grad = False
        self.encoder.to(dev).eval()
        self.recnet.to(dev)
        self._flat, self._flat_bound = None, False
        self._buckets = None
        if self.isTrain:
            self.recnet.train()
            self._order_parameters()
            params = [p for p in self.recnet.parameters() if p.requires_grad]
            if opts.optimizer.lower() != "adam":
                raise NotImplementedError("run.py uses Adam (run.py:11)")
            # clip_grad_value_(1.0) + Adam fused into one kernel launch over all RecNet/head tensors
            self.optim = FusedClipAdam(self._ordered_params, opts.lr, betas=(opts.beta1, opts.beta2),
                                       weight_decay=opts.weight_decay, clip_value=1.0)
            self.sch = optim.lr_scheduler.MultiStepLR(self.optim, [5000, 10000, 15000], gamma=0.5)
            if dev.type == "cuda":
                self.bind_flat_gradients()         # every .grad is a view of one flat buffer from the start
        else:
            self.recnet.eval()
        # literal=True keeps the reference's op sequence for the losses (ATen under autograd over the public 7-tuples);
        # the default runs the fused loss kernels on the batched engine
        self.literal = bool(getattr(opts, "literal", False))
        self.merge_encoder_batches = bool(getattr(opts, "merge_encoder_batches", True))
        self.overlap_allreduce = bool(getattr(opts, "overlap_allreduce", True))
        self.mse_loss = nn.MSELoss()
        self.triplet = TripletLoss()
        self.cross_entropy = nn.CrossEntropyLoss()
        self._graph = None
        self._static = None
        self._lws = {}
        self._comm_stream = None
        self._ar_events = []
        if not self.isTrain or getattr(opts, "continue_train", False):   # reference trainer.py:89-90
            which = getattr(opts, "which_file", None)
            if which:
                self.load_model(which)

    def clone_model(self):
        """Reference trainer.py:97-113: fresh copies of both networks carrying the current weights."""
        net_copy = {"Senet": Backbone(num_layers=50, drop_ratio=0.6, mode="ir_se"),
                    "Recnet": RecNet(norm_type=self.norm_type, relu_type=self.relu_type)}
        net_copy["Senet"].load_state_dict(self.encoder.state_dict())
        net_copy["Recnet"].load_state_dict(self.recnet.state_dict())
        for p in net_copy["Senet"].parameters():
            p.requires_grad = False
        net_copy["Senet"].to(self.opts.device)
        net_copy["Recnet"].to(self.opts.device)
        return net_copy

    # ------------------------------------------------------------------------------------------------------
    def _order_parameters(self):
        """Parameters grouped by the order in which the backward pass completes their gradients: the flat gradient
        buffer is laid out bucket after bucket so each bucket is one contiguous range for the all-reduce."""
        named = [(k, p) for k, p in self.recnet.named_parameters() if p.requires_grad]
        order = []
        for prefix in _BUCKET_ORDER:
            order += [(k, p) for k, p in named if k.startswith(prefix)]
        rest = [(k, p) for k, p in named if not any(k.startswith(pre) for pre in _BUCKET_ORDER)]
        self._named_ordered = order + rest
        self._ordered_params = [p for _, p in self._named_ordered]

    def step(self, img1, img2, label):
        """One whole iteration. After capture_step() the iteration (backbone forward over both image sets, batched RecNet
        forward / backward, losses, [bucketed all-reduce], clip + Adam) replays from ONE CUDA graph; the learning
        rate and the Adam step count live on the device."""
        if self._graph is None:
            self.set_input(img1, img2, label)
            self.forward()
            self.optimizer_parameters(0)
        else:
            for dst, src in zip(self._static, (img1, img2, label)):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
            # the replay updated parameters and BatchNorm buffers through raw pointers: invalidate packed-weight caches
            _lib.bump_weights_generation()
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
        for p in self._ordered_params:
            ok = ok and p.grad is not None and p.grad.data_ptr() == base + 4 * off
            off += p.numel()
        return ok

    def bind_flat_gradients(self):
        """Make every RecNet/head .grad a view into ONE flat fp32 buffer: the data-parallel exchange needs no pack /
        unpack copies and the backward kernels write the gradients in place."""
        params = self._ordered_params
        n = sum(p.numel() for p in params)
        self._flat = torch.zeros(n, dtype=torch.float32, device=params[0].device)
        off = 0
        self._buckets = []
        cur = None
        for (k, p) in self._named_ordered:
            p.grad = self._flat[off:off + p.numel()].view_as(p)
            pre = next((b for b in _BUCKET_ORDER if k.startswith(b)), "other")
            if cur is None or cur[0] != pre:
                cur = [pre, off, off]
                self._buckets.append(cur)
            off += p.numel()
            cur[2] = off
        self._flat_bound = True

    def capture_step(self, img1, img2, label, warmup=3):
        """Record the iteration into one CUDA graph (the NCCL all-reduces of the gradient buckets are captured on their
        side stream, fork / join by events inside the graph). The warm-up iterations are real optimizer steps on the
        capture batch: parameters, Adam state and BatchNorm buffers are restored afterwards, so capturing has no side
        effect on the training trajectory."""
        if not self._flat_bound:
            self.bind_flat_gradients()
        self._static = (img1.clone(), img2.clone(), label.clone())
        snap_p = [p.detach().clone() for p in self.recnet.parameters()]
        snap_b = [b.detach().clone() for b in self.recnet.buffers()]
        snap_o = self.optim.snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self.set_input(*self._static)
                self.forward()
                self.optimizer_parameters(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.set_input(*self._static)
            self.forward()
            self.optimizer_parameters(0)
        with torch.no_grad():
            for p, s in zip(self.recnet.parameters(), snap_p):
                p.copy_(s)
            for b, s in zip(self.recnet.buffers(), snap_b):
                b.copy_(s)
        self.optim.restore(snap_o)
        _lib.bump_weights_generation()
        self._graph = graph

    def set_input(self, img1, img2, label):
        self.nonocl, self.ocl, self.gt_label = img1, img2, label

    # ------------------------------------------------------------------------------------------------------
    def forward(self):
        with torch.no_grad():
            n = self.nonocl.shape[0]
            if self.merge_encoder_batches and self.nonocl.shape == self.ocl.shape:
                # the frozen eval-mode backbone is per-image: one forward over both image sets (better tile / wave
                # occupancy than two half-size ones), then split — same values as the two calls of trainer.py:141-142
                y, f = self.encoder(torch.cat((self.nonocl, self.ocl), 0))
            else:
                y1, f1 = self.encoder(self.nonocl)
                y2, f2 = self.encoder(self.ocl)
                y, f = torch.cat((y1, y2), 0), torch.cat((f1, f2), 0)
            self.feat_map_non, self.feat_map_ocl = y[:n], y[n:]
            self.feat_extract_non, self.feat_extract_ocl = f[:n], f[n:]
            self._y, self._f = y, f
        if self.literal or not self.recnet.training:
            return self._forward_literal()
        self._forward_engine()

    def _forward_literal(self):
        """The reference's own sequence: two public RecNet calls returning 7-tuples (trainer.py:144-152)."""
        (self.f_non, self.pred_loss_non, self.pred_label_non, self.M_space_non, self.M_channel_non, self.space_non,
         self.channel_non) = self.recnet(self.feat_map_non, self.gt_label)
        (self.f_ocl, self.pred_loss_ocl, self.pred_label_ocl, self.M_space_ocl, self.M_channel_ocl, self.space_ocl,
         self.channel_ocl) = self.recnet(self.feat_map_ocl, self.gt_label)
        pred = self.pred_label_ocl.detach().argmax(1)
        self.pred_label = pred
        self._correct = pred.eq(self.gt_label).sum()

    def _loss_ws(self, n, dev):
        key = (n, str(dev))
        lw = self._lws.get(key)
        if lw is not None:
            return lw
        lw = losses.LossWorkspace(n, dev)
        f32 = dict(dtype=torch.float32, device=dev)
        lw.v = torch.zeros(2 * n, 512, **f32)
        lw.dv = torch.zeros(2 * n, 512, **f32)
        lw.ce = torch.zeros(2, **f32)
        w3 = float(self.opts.loss_weight[3])
        lw.gloss = torch.tensor([w3 / (1e-8 + w3), w3], **f32)       # d total / d CE_non, d total / d CE_ocl (:174,:178)
        lw.head = head.GroupedHead(self.recnet.classifier, 2 * n, n, dev)
        self._lws[key] = lw
        return lw

    def _forward_engine(self):
        eng = recnet_train.engine(self.recnet)
        n = self.nonocl.shape[0]
        dev = self._y.device
        lw = self._loss_ws(n, dev)
        self._lw = lw
        if n % 32 == 0:
            self._calls = [(eng.forward(self._y, 2, slot=0, v_out=lw.v), 0, 2)]
        else:       # BatchNorm statistics per call need group-aligned tiles: run the two calls one after the other
            self._calls = [(eng.forward(self._y[:n].contiguous(), 1, slot=0, v_out=lw.v[:n]), 0, 1),
                           (eng.forward(self._y[n:].contiguous(), 1, slot=1, v_out=lw.v[n:]), 1, 1)]
        self.f_non, self.f_ocl = lw.v[:n], lw.v[n:]
        lw.head.forward(lw.v, self.gt_label, lw.ce)
        self.pred_label = lw.head.pred[n:]
        self._correct = self.pred_label.eq(self.gt_label).sum()          # no host sync here; .item() in get_current_values

    # ------------------------------------------------------------------------------------------------------
    def backward(self):
        if self.literal or not self.recnet.training:
            return self._backward_literal()
        lib = _lib.load()
        st = _lib.stream_ptr()
        lw = self._lw
        n = self.nonocl.shape[0]
        w = [float(v) for v in self.opts.loss_weight]
        x_non = self.feat_map_non
        grads = {k: p.grad for k, p in self.recnet.named_parameters()}
        for ws, g0, G in self._calls:
            losses.selfsim_channel(lw, ws.fc, x_non, G * n, n, g0, w[0])        # trainer.py:157-165
            losses.selfsim_space(lw, ws.fs, x_non, G * n, n, g0, w[0])
        losses.triplet_identity(lw, lw.v[:n], lw.v[n:], self.feat_extract_non, self.feat_extract_ocl, w[1], w[2])   # :167-171
        dvh = lw.head.backward(lw.gloss, grads["classifier.weight"])            # :173-176
        _lib.check(lib.ffr_add3_f32(_P(dvh), _P(lw.dvl), None, _P(lw.dv), 2 * n * 512, st), "add3")
        losses.finalize(lw, lw.ce, w)
        self.loss_items = [lw.out[i] for i in range(4)]
        self.pos_loss, self.neg_loss = lw.out[4], lw.out[5]
        self._bucket_ready("classifier")
        eng = recnet_train.engine(self.recnet)
        for i, (ws, g0, G) in enumerate(self._calls):
            r0 = g0 * n
            last = i == len(self._calls) - 1
            eng.backward(ws, grads, dv=lw.dv[r0:r0 + G * n], dfs=lw.dfs[r0 * 81:(r0 + G * n) * 81],
                         dfc=lw.dfc[r0 * 81:(r0 + G * n) * 81], accumulate=(i > 0),
                         on_stage=(self._bucket_ready if last else None))

    def _backward_literal(self):
        # trainer.py:157-161 calls selfSimilarity five times and discards one of the two Grams in four of them
        ss_space, ss_channel = selfSimilarity(self.feat_map_non)
        ss_space_non, _ = selfSimilarity(self.space_non)
        ss_space_ocl, _ = selfSimilarity(self.space_ocl)
        _, ss_channel_non = selfSimilarity(self.channel_non)
        _, ss_channel_ocl = selfSimilarity(self.channel_ocl)
        mse = self.mse_loss
        l_space = (mse(ss_space, ss_space_non) + mse(ss_space, ss_space_ocl)) / 2
        l_channel = (mse(ss_channel, ss_channel_non) + mse(ss_channel, ss_channel_ocl)) / 2
        items = [(l_space + l_channel) / 2]
        t, self.pos_loss, self.neg_loss = self.triplet(self.f_ocl, self.feat_extract_non, self.feat_extract_ocl)
        items.append(t)
        items.append((mse(self.f_non, self.feat_extract_non) + mse(self.f_ocl, self.feat_extract_non)) / 2)
        ce = lambda logits: self.cross_entropy(logits, self.gt_label)
        items.append(ce(self.pred_loss_non) / (1e-8 + self.opts.loss_weight[3]) + ce(self.pred_loss_ocl))
        self.loss_items = [l * w for l, w in zip(items, self.opts.loss_weight)]
        sum(self.loss_items).backward()

    # ------------------------------------------------------------------------------------------------------
    def _dp(self):
        import torch.distributed as dist
        if not getattr(self.opts, "data_parallel", True) if hasattr(self, "opts") else False:
            return False
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    @staticmethod
    def _allreduce_avg(t):
        """In-place mean over ranks (NCCL: ReduceOp.AVG inside the collective; gloo has no AVG: sum, then divide)."""
        import torch.distributed as dist
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.div_(dist.get_world_size())

    def _bucket_ready(self, prefix):
        """Called by the backward pass when every gradient of a bucket has been written: launch its all-reduce (mean) on
        the communication stream, ordered after the producing kernels by an event (CPU tensors: synchronously)."""
        if not (self._dp() and getattr(self, "overlap_allreduce", True) and self._flat_bound):
            return
        cuda = self._flat.is_cuda
        if cuda:
            main = torch.cuda.current_stream()
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream()
            ev = torch.cuda.Event()
            ev.record(main)
        for name, lo, hi in self._buckets:
            if name != prefix:
                continue
            if not cuda:
                self._allreduce_avg(self._flat[lo:hi])
                self._ar_events.append(None)
                continue
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._allreduce_avg(self._flat[lo:hi])
                done = torch.cuda.Event()
                done.record(self._comm_stream)
            self._ar_events.append(done)

    def allreduce_gradients(self):
        """Average the RecNet/head gradients over ranks. With overlap the buckets were launched during backward():
        only the join remains; otherwise one all-reduce of the whole flat buffer (or a pack / unpack when the gradients
        are not views of it)."""
        if not self._dp():
            return
        if self._ar_events:
            for ev in self._ar_events:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
            self._ar_events = []
            return
        if self._flat_bound and self._grads_bound():
            self._allreduce_avg(self._flat)
            return
        params = [p for p in self.recnet.parameters() if p.grad is not None]
        n = sum(p.numel() for p in params)
        flat = torch.empty(n, dtype=torch.float32, device=params[0].device)
        off = 0
        for p in params:
            flat[off:off + p.numel()].copy_(p.grad.reshape(-1))
            off += p.numel()
        self._allreduce_avg(flat)
        off = 0
        for p in params:
            p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
            off += p.numel()

    def optimizer_parameters(self, cur_iters=0):
        if self.literal:
            self.zero_grad()                           # autograd accumulates; the engine path overwrites every gradient
        self.backward()
        self.allreduce_gradients()                     # before clipping, as reduce-then-clip in the reference
        self.optim.step()                              # clip_grad_value_(1.0) + Adam, one fused launch

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

    def get_current_values(self):
        keys = ["SelfSimilarityLoss", "TripletLoss", "IdentityLoss", "ClassifierLoss"]
        d = {k: "{:.4f}".format(float(v.detach())) for k, v in zip(keys, self.loss_items)}
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
