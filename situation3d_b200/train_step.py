"""SQA3D-shaped training step of the backbone (BASELINE.json config 4).

Training mode runs the drop-in modules operator by operator (autograd through the ``*_grad`` kernels,
train-mode BatchNorm with per-replica statistics, as in the reference).  Multi-GPU: one process per
GPU, scenes sharded by rank, and ONE gradient all-reduce per step over a flat bucket
(``sharding.FlatGradAllReduce``), launched on a side stream as soon as backward has finished so that
it overlaps with the host-side bookkeeping of the step; the optimizer waits for it.
"""
import torch
import torch.distributed as dist

from .sharding import FlatGradAllReduce


class BackboneTrainer:
    def __init__(self, net, lr=1e-3, group=None):
        self.net = net.train()
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.bucket = FlatGradAllReduce(net, group)
        self.opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=0.9)
        self.comm_stream = torch.cuda.Stream() if self.distributed else None
        self.group = group

    def step(self, point_clouds):
        """One step on this rank's scenes; returns the (local) loss tensor."""
        self.bucket.zero()
        out = self.net({"point_clouds": point_clouds})
        loss = out["fp2_features"].square().mean()
        loss.backward()                                   # gradients land in the flat bucket
        if self.distributed:
            main = torch.cuda.current_stream()
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.bucket.flat.div_(dist.get_world_size(self.group))
            main.wait_stream(self.comm_stream)
        self.opt.step()
        return loss.detach()
