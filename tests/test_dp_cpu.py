"""CPU test (gloo, world_size 2) of the data-parallel host logic: the flat-bucket gradient all-reduce of
ffr_net_b200.trainer.Trainer averages gradients across ranks and leaves every tensor's shape/values consistent."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ffr_net_b200.trainer import Trainer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.PReLU(5), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        p.grad = torch.randn(p.shape, generator=g)
    local = [p.grad.clone() for p in net.parameters()]
    t = Trainer.__new__(Trainer)          # only the all-reduce helper is under test (no GPU needed)
    t.recnet, t._flat, t._flat_bound, t._ar_events = net, None, False, []
    t.allreduce_gradients()
    # expected: mean over ranks of the per-rank gradients
    exp = []
    for i, p in enumerate(net.parameters()):
        parts = [torch.zeros_like(local[i]) for _ in range(world)]
        dist.all_gather(parts, local[i])
        exp.append(sum(parts) / world)
    ok = all(torch.allclose(p.grad, e, atol=1e-6) for p, e in zip(net.parameters(), exp))
    # flat-bound gradients (what capture_step uses under DP): .grad tensors are views of one buffer, autograd
    # accumulates into them in place, and the exchange is a single in-place all-reduce with no copies
    t2 = Trainer.__new__(Trainer)
    net2 = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.PReLU(5), torch.nn.Linear(5, 3))
    net2.load_state_dict(net.state_dict())
    t2.recnet, t2._flat, t2._flat_bound, t2._ar_events = net2, None, False, []
    t2._order_parameters()
    t2.bind_flat_gradients()
    x = torch.randn(4, 7, generator=g)
    net2(x).square().sum().backward()
    base = t2._flat.data_ptr()
    ok = ok and all(p.grad.data_ptr() >= base and p.grad.data_ptr() < base + 4 * t2._flat.numel()
                    for p in net2.parameters())
    local2 = [p.grad.clone() for p in net2.parameters()]
    ok = ok and all(float(l.abs().sum()) > 0 for l in local2)
    t2.allreduce_gradients()
    for i, p in enumerate(net2.parameters()):
        parts = [torch.zeros_like(local2[i]) for _ in range(world)]
        dist.all_gather(parts, local2[i])
        ok = ok and torch.allclose(p.grad, sum(parts) / world, atol=1e-6)
    # bucketed exchange (what backward() triggers bucket by bucket under DP): every bucket is averaged exactly once and
    # allreduce_gradients() afterwards only joins
    t3 = Trainer.__new__(Trainer)
    net3 = torch.nn.ModuleDict({"classifier": torch.nn.Linear(4, 3), "Conv4Merge": torch.nn.Linear(3, 2),
                                "Conv4Space": torch.nn.PReLU(2)})
    t3.recnet, t3._flat, t3._flat_bound, t3._ar_events, t3._comm_stream = net3, None, False, [], None
    t3.overlap_allreduce = True
    t3._order_parameters()
    t3.bind_flat_gradients()
    names = [b[0] for b in t3._buckets]
    ok = ok and names == ["classifier", "Conv4Merge", "Conv4Space"]
    ok = ok and t3._buckets[0][1] == 0 and t3._buckets[-1][2] == t3._flat.numel()
    with torch.no_grad():
        t3._flat.copy_(torch.randn(t3._flat.numel(), generator=g))
    local3 = t3._flat.clone()
    for nme in names:
        t3._bucket_ready(nme)
    ok = ok and len(t3._ar_events) == 3
    t3.allreduce_gradients()
    parts = [torch.zeros_like(local3) for _ in range(world)]
    dist.all_gather(parts, local3)
    ok = ok and torch.allclose(t3._flat, sum(parts) / world, atol=1e-6) and t3._ar_events == []
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
