"""SQA3D-shaped training step of the backbone (BASELINE.json config 4).

Training mode runs the backbone on channel-last rows (``train_rows.py``: this library's sampling / ball-query / 3-NN
kernels for the indices, row gathers, one GEMM per 1x1 convolution, train-mode BatchNorm with per-replica statistics
as in the reference; ``net.train_layout = "reference"`` selects the operator-by-operator wiring of the reference).
Multi-GPU: one process per GPU, scenes sharded by rank, and the gradient all-reduce over a flat bucket
(``sharding.FlatGradAllReduce``) split into three chunks that are reduced on a side stream WHILE backward is still
running -- FP layers first, SA1/SA2 last -- so that only the last chunk's collective is exposed; the optimizer waits
for it.  BatchNorm statistics are not synchronised (the reference has no SyncBN).
"""
import torch
import torch.distributed as dist

from .sharding import FlatGradAllReduce


class BackboneTrainer:
    def __init__(self, net, lr=1e-3, group=None, overlap=True):
        self.net = net.train()
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.bucket = FlatGradAllReduce(net, group)
        self.opt = torch.optim.SGD(net.parameters(), lr=lr, momentum=0.9)
        self.comm_stream = torch.cuda.Stream() if (self.distributed and next(net.parameters()).is_cuda) else None
        self.group = group
        self.overlap = overlap and self.distributed
        if self.overlap:
            self.bucket.enable_overlap(3, self.comm_stream)

    def step(self, point_clouds, next_point_clouds=None):
        """One step on this rank's scenes; returns the (local) loss tensor.  ``next_point_clouds`` (the batch the data
        loader already holds for the following step) has its sampling chain started under this step's backward."""
        self.bucket.zero()
        out = self.net({"point_clouds": point_clouds})
        loss = out["fp2_features"].square().mean()
        if next_point_clouds is not None and hasattr(self.net, "prefetch_sampling"):
            self.net.prefetch_sampling(next_point_clouds)
        loss.backward()                                   # gradients land in the flat bucket; chunk hooks start the collectives
        if self.overlap:
            self.bucket.finish()
        elif self.distributed:
            main = torch.cuda.current_stream()
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.bucket.flat, op=dist.ReduceOp.SUM, group=self.group)
                self.bucket.flat.div_(dist.get_world_size(self.group))
            main.wait_stream(self.comm_stream)
        self.opt.step()
        return loss.detach()
