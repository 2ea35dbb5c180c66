"""Worker of tests/test_dp_gpu.py (launched with torchrun, one process per GPU, NCCL): the data-parallel training step.
Every rank computes the gradients of its own shard (BatchNorm statistics per rank, like the reference's per-replica
data_parallel), the bucketed all-reduce averages them during backward; the result must equal the average of the per-shard
gradients computed on ONE GPU, and after clip + Adam every rank must hold identical parameters."""
import faulthandler
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ffr_net_b200 import synth                               # noqa: E402
from ffr_net_b200.recnet import RecNet                       # noqa: E402
from ffr_net_b200.trainer import Trainer, default_opts       # noqa: E402


def shard(rank, n):
    a = synth.synth_faces(n, seed=100 + rank)
    b = synth.synth_faces(n, seed=100 + rank, masked=True)
    label = torch.randint(0, 10575, (n,), generator=torch.Generator().manual_seed(100 + rank))
    return a.cuda(), b.cuda(), label.cuda()


def grads_of(tr, data):
    tr.set_input(*data)
    tr.forward()
    tr.zero_grad()
    tr.backward()
    tr.allreduce_gradients()
    torch.cuda.synchronize()
    return {k: p.grad.clone() for k, p in tr.recnet.named_parameters()}


T0 = time.time()


def note(rank, msg):
    print("[dp_worker rank %d %6.1fs] %s" % (rank, time.time() - T0, msg), file=sys.stderr, flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    faulthandler.dump_traceback_later(int(os.environ.get("DP_WORKER_DUMP_AFTER", "150")), exit=True)   # a hang names its line
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 32
    bsd, rsd = synth.synth_backbone_state_dict(0), synth.synth_recnet_state_dict(0)

    def make(**kw):
        rec = RecNet()
        rec.load_state_dict(rsd)
        return Trainer(default_opts(lr=1e-3, **kw), recnet=rec, encoder_weights=bsd)
    out = {}
    # (1) data-parallel gradients: bucketed / overlapped exchange, and the single flat all-reduce
    note(rank, "process group up")
    g_overlap = grads_of(make(overlap_allreduce=True), shard(rank, n))
    note(rank, "overlapped exchange done")
    g_flat = grads_of(make(overlap_allreduce=False), shard(rank, n))
    note(rank, "flat exchange done")
    out["overlap_equals_flat"] = all(torch.equal(g_overlap[k], g_flat[k]) for k in g_flat)
    # (2) expected: average of the per-shard gradients, all computed on this GPU without any exchange
    per_shard = [grads_of(make(data_parallel=False), shard(r, n)) for r in range(world)]
    def errors(got):
        e = {}
        for k in got:
            exp = sum(g[k].double() for g in per_shard) / world
            e[k] = ((got[k].double() - exp).norm() / (exp.norm() + 1e-30)).item()
        return e
    e_flat, e_overlap = errors(g_flat), errors(g_overlap)
    out["worst_rel_err_vs_single_gpu_average"] = max(e_flat.values())
    out["worst_rel_err_overlap_vs_single_gpu_average"] = max(e_overlap.values())
    out["worst_params_flat"] = sorted(e_flat.items(), key=lambda kv: -kv[1])[:4]
    out["worst_params_overlap"] = sorted(e_overlap.items(), key=lambda kv: -kv[1])[:4]
    # the rank's own shard, computed twice by two trainer instances, must be bit-identical (deterministic engine)
    again = grads_of(make(data_parallel=False), shard(rank, n))
    out["own_shard_reproducible"] = all(torch.equal(again[k], per_shard[rank][k]) for k in again)
    note(rank, "single-GPU shard gradients done")
    # (3) a captured step under DP: every rank ends with identical parameters
    tr = make()
    data = shard(rank, n)
    tr.capture_step(*data, warmup=2)
    for _ in range(3):
        tr.step(*data)
    torch.cuda.synchronize()
    note(rank, "captured steps done")
    flat = torch.cat([p.detach().reshape(-1) for p in tr.recnet.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    out["params_identical_across_ranks"] = all(torch.equal(parts[0], q) for q in parts)
    out["finite"] = bool(torch.isfinite(flat).all())
    if rank == 0:
        print("DP_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    del tr                        # the captured graph holds NCCL kernels: destroying the process group under it hangs
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
