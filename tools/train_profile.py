"""torch.profiler breakdown of one training step (GPU box): which kernels / ops dominate. Writes gpurun_out/train_profile.txt."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ffr_net_b200 import synth as ob
from ffr_net_b200 import synth as orr
from ffr_net_b200.backbone import Backbone
from ffr_net_b200.recnet import RecNet
from ffr_net_b200.trainer import Trainer, default_opts

dev = torch.device("cuda")
enc = Backbone(50, 0.6, "ir_se")
enc.load_state_dict(ob.synth_backbone_state_dict(0))
rec = RecNet()
rec.load_state_dict(orr.synth_recnet_state_dict(0))
tr = Trainer(default_opts(lr=1e-4), encoder=enc, recnet=rec)
pairs = 256
a = ob.synth_faces(64, seed=1).repeat(4, 1, 1, 1).to(dev)
b = ob.synth_faces(64, seed=1, masked=True).repeat(4, 1, 1, 1).to(dev)
label = torch.randint(0, 10575, (pairs,)).to(dev)


def step():
    tr.set_input(a, b, label)
    tr.forward()
    tr.optimizer_parameters(0)
    tr.update_learning_rate()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70)
open("gpurun_out/train_profile.txt", "w").write(txt)
print(txt)
# kernels only (device-side events), with their share of the step
from torch.autograd import DeviceType
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
ks.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ks)
lines = ["# kernels of one training step (256 pairs), device time %.3f ms over %d launches" %
         (tot / 1e3, sum(e.count for e in ks)), "share,total_us,launches,kernel"]
for e in ks[:60]:
    lines.append("%.3f,%.1f,%d,%s" % (e.self_device_time_total / tot, e.self_device_time_total, e.count, e.key[:110]))
open("gpurun_out/train_profile_kernels.csv", "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
shp = prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=40)
open("gpurun_out/train_profile_shapes.txt", "w").write(shp)
import time
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
