"""A/B of the gradient exchange inside the data-parallel training step (torchrun, one rank per GPU): bucketed all-reduces
overlapped with the backward pass on a communication stream vs one flat all-reduce after it, both as ONE CUDA graph.
NCCL_MAX_CTAS (environment) bounds the SMs NCCL may hold while the persistent tcgen05 GEMMs of the backward run.
  torchrun --nproc-per-node 2 tools/dp_overlap_ab.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet
from ffr_net_b200.trainer import Trainer, default_opts

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pairs = 256
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(ob.synth_backbone_state_dict(0))
enc = enc.to(dev).eval()
a = ob.synth_faces(64, seed=10 + rank).repeat(pairs // 64, 1, 1, 1).to(dev)
b = ob.synth_faces(64, seed=10 + rank, masked=True).repeat(pairs // 64, 1, 1, 1).to(dev)
label = torch.randint(0, 10575, (pairs,), generator=torch.Generator().manual_seed(rank)).to(dev)
trainers = {}
for overlap in (True, False):
    rec = RecNet()
    rec.load_state_dict(ob.synth_recnet_state_dict(0))
    tr = Trainer(default_opts(lr=1e-4, device=str(dev), overlap_allreduce=overlap), encoder=enc, recnet=rec)
    tr.capture_step(a, b, label, warmup=3)
    trainers[overlap] = tr
res = {True: [], False: []}
for rnd in range(6):
    for overlap in ((True, False) if rnd % 2 == 0 else (False, True)):
        tr = trainers[overlap]
        for _ in range(2):
            tr.step(a, b, label)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            tr.step(a, b, label)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 8], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[overlap].append(round(float(t), 3))
if rank == 0:
    out = {"world": world, "NCCL_MAX_CTAS": os.environ.get("NCCL_MAX_CTAS"), "overlapped_buckets_ms": res[True], "flat_after_backward_ms": res[False],
           "median_overlapped": sorted(res[True])[len(res[True]) // 2], "median_flat": sorted(res[False])[len(res[False]) // 2]}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/dp_overlap_ab.jsonl", "a") as f:
        f.write(json.dumps(out) + "\n")
# Trainers holding CUDA graphs with captured NCCL collectives: tearing the communicator down at interpreter exit can block
# for minutes (seen: 180 s per run); leave without running destructors, like tests/dp_worker.py
sys.stdout.flush()
os._exit(0)
