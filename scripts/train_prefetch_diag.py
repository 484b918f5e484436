"""Per-step times of the training step with and without the sampling prefetch (GPU box): is the occasional slow
measurement of the prefetched variant uniform or spiky?   python scripts/train_prefetch_diag.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch
from situation3d_b200.train_step import BackboneTrainer
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)
dev = torch.device("cuda", 0)
# optional: a first phase that leaves the allocator in the state the full bench has when config 4 starts
if os.environ.get("DIRTY", "0") == "1":
    junk = [torch.empty(int(2 ** (20 + (i % 9))), device=dev) for i in range(400)]
    del junk
net = Pointnet2Backbone(input_feature_dim=129).to(dev)
tr = BackboneTrainer(net)
pcs = [torch.from_numpy(make_batch(8, 40000, 129, first_seed=100 * j)).to(dev) for j in range(2)]
for prefetch in (False, True, False, True):
    for i in range(3):
        tr.step(pcs[i % 2], pcs[(i + 1) % 2] if prefetch else None)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
    t0 = time.time()
    ev[0].record()
    for i in range(24):
        tr.step(pcs[(i + 1) % 2], pcs[i % 2] if prefetch else None)
        ev[i + 1].record()
    host = (time.time() - t0) / 24 * 1e3
    torch.cuda.synchronize()
    dt = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(24)])
    print("prefetch=%s: mean %.2f ms, median %.2f, min %.2f, max %.2f, host enqueue %.2f ms/step  | %s" % (
        prefetch, dt.mean(), np.median(dt), dt.min(), dt.max(), host, " ".join("%.1f" % v for v in dt[:12])), flush=True)
