import os, sys, collections
import torch
from torch.profiler import ProfilerActivity, profile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone as ob
from oracle import recnet as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet
from ffr_net_b200.trainer import Trainer, default_opts
dev = torch.device("cuda")
enc = Backbone(50, 0.6, "ir_se"); enc.load_state_dict(ob.synth_backbone_state_dict(0))
rec = RecNet(); rec.load_state_dict(orr.synth_recnet_state_dict(0))
tr = Trainer(default_opts(lr=1e-4), encoder=enc, recnet=rec)
pairs = 256
a = ob.synth_faces(64, seed=1).repeat(4, 1, 1, 1).to(dev)
b = ob.synth_faces(64, seed=1, masked=True).repeat(4, 1, 1, 1).to(dev)
label = torch.randint(0, 10575, (pairs,)).to(dev)
import ffr_net_b200.recnet_train as rt


def count(tag):
    for _ in range(2):
        tr.step(a, b, label)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU], record_shapes=True) as prof:
        tr.step(a, b, label)
        torch.cuda.synchronize()
    c = collections.Counter()
    for e in prof.events():
        if e.name.startswith("aten::") and e.input_shapes and e.input_shapes[0] == [256]:
            c[e.name] += 1
    tot = collections.Counter(e.name for e in prof.events() if e.name.startswith("aten::"))
    print(tag, "ops on [256]:", dict(c), "| total aten ops:", sum(tot.values()), "| top:", tot.most_common(6))


for _ in range(2):
    tr.step(a, b, label)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.step(a, b, label)
    torch.cuda.synchronize()
par = collections.Counter()
for e in prof.events():
    if e.name in ("aten::add", "aten::select") and e.input_shapes:
        chain, q = [], e.cpu_parent
        while q is not None and len(chain) < 4:
            chain.append(q.name[:50]); q = q.cpu_parent
        par[(e.name, str(e.input_shapes[0]), " < ".join(chain))] += 1
for k, v in par.most_common(10):
    print(v, k)
