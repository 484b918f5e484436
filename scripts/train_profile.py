"""Kernel-time breakdown of the config-4 training step (BackboneTrainer, 8 scenes x 40k points) with torch.profiler
(CUPTI): which kernels the 16 ms go to.  GPU box.   python scripts/train_profile.py [rows|reference]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch
from situation3d_b200.train_step import BackboneTrainer
from torch.profiler import profile, ProfilerActivity

layout = sys.argv[1] if len(sys.argv) > 1 else "rows"
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
net = Pointnet2Backbone(input_feature_dim=129).to(dev)
if layout != "rows":
    net.train_layout = "reference"
tr = BackboneTrainer(net)
pc = torch.from_numpy(make_batch(8, 40000, 129)).to(dev)
for _ in range(3):
    tr.step(pc)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    tr.step(pc)
e.record(); torch.cuda.synchronize()
print("ms per step (%s): %.3f" % (layout, s.elapsed_time(e) / 5))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step(pc)
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages():
    t = getattr(ev, "device_time_total", None)
    if t is None:
        t = getattr(ev, "cuda_time_total", 0)
    if ev.device_type.name == "CUDA" or (t and ev.self_device_time_total > 0 if hasattr(ev, "self_device_time_total") else False):
        rows.append((ev.self_device_time_total / 3.0 if hasattr(ev, "self_device_time_total") else t / 3.0, ev.count // 3, ev.key[:110]))
rows = [r for r in rows if r[0] > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("sum of kernel time per step: %.3f ms over %d kernel kinds" % (tot / 1e3, len(rows)))
for t, n, k in rows[:40]:
    print("%9.1f us %5.1f%% x%-4d %s" % (t, 100 * t / tot, n, k))
