"""Config 4: backbone forward+backward with one NCCL gradient all-reduce per step (run under torchrun)."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.synthetic import make_batch
from situation3d_b200.train_step import BackboneTrainer

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, steps = int(os.environ.get("B", 8)), int(os.environ.get("STEPS", 5))
torch.manual_seed(0)
net = Pointnet2Backbone(input_feature_dim=129).cuda()
tr = BackboneTrainer(net)
pc = torch.from_numpy(make_batch(B, 40000, 129, first_seed=rank * B)).cuda()
for _ in range(2):
    tr.step(pc)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(steps):
    loss = tr.step(pc)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
# exposed all-reduce time: the same collective timed alone
t = torch.zeros_like(tr.bucket.flat)
if world > 1:
    dist.all_reduce(t); torch.cuda.synchronize()
    s.record()
    for _ in range(20):
        dist.all_reduce(t)
    e.record(); torch.cuda.synchronize()
    ar_us = 1e3 * s.elapsed_time(e) / 20
else:
    ar_us = 0.0
if rank == 0:
    print(json.dumps({"config": "training step, B=%d/GPU, 40k points, train-mode BN, unfused operators + autograd" % B,
                      "n_gpus": world, "ms_per_step": ms, "scenes_per_s": world * B * 1e3 / ms, "loss": float(loss),
                      "grad_bytes": tr.bucket.flat.numel() * 4, "allreduce_us_alone": ar_us}))
if world > 1:
    dist.destroy_process_group()
