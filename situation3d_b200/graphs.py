"""CUDA-graph replay of the fused backbone step.

A ``Pointnet2Backbone`` forward is 23 kernels of this library plus a handful of tensor allocations, issued from
Python over two streams (the sampling chain runs ahead on a side stream).  Enqueuing that costs ~0.8 ms of host
time per step -- as much as the GPU needs for the step once several batches are in flight -- so callers with a
fixed input shape capture the step once and replay it:

    step = GraphedBackbone(net, example_pc)       # captures on its own stream
    out = step(pc)                                # copy into the static input, one cudaGraphLaunch
    step.stream.synchronize()

The graph holds both streams' work (fork / join through events), its tensors live in the graph's private memory
pool, and ``out`` is the same dict of static output tensors on every call.  The capture also freezes the weight
images of the fused layers: capture again after the module's parameters change.
"""
import torch

from .streams import lane_stream


def _as_inputs(example):
    """A (B,N,3+C) f32 tensor -> {"point_clouds": t}; an (xyz f32, feature rows bf16) pair -- the compact transport
    format of ``Pointnet2Backbone.pack_point_clouds`` -- -> {"xyz": ..., "feature_rows_bf16": ...}; dicts pass."""
    if isinstance(example, dict):
        return dict(example)
    if isinstance(example, (tuple, list)):
        return {"xyz": example[0], "feature_rows_bf16": example[1]}
    return {"point_clouds": example}


def pick_sampling_mode(scenes_in_flight):
    """"throughput" once the caller keeps enough scenes in flight to fill the GPU with one sampling CTA per scene
    (fused.sampling_mode, include/pn2_b200.h PN2_FPS_THROUGHPUT), "latency" otherwise."""
    return "throughput" if scenes_in_flight >= 128 else "latency"


class GraphedBackbone:
    def __init__(self, net, example, stream=None, static_input=None, warmup=2, sampling_mode="latency"):
        """``sampling_mode``: which FPS kernel the captured step uses for mid-sized scenes (results are identical):
        "latency" = fastest single step, "throughput" = least SM time, for many steps in flight (pick_sampling_mode)."""
        assert not net.training, "the fused (forward-only) path is what gets captured"
        self.net = net
        example = _as_inputs(example)
        first = next(iter(example.values()))
        self.stream = stream if stream is not None else lane_stream(first.device)
        self.static_in = _as_inputs(static_input) if static_input is not None else {k: torch.empty_like(v) for k, v in example.items()}
        from ._lib import lib
        from . import fused
        with torch.no_grad(), fused.sampling_mode(sampling_mode):
            with torch.cuda.stream(self.stream):
                if static_input is None:
                    for k, v in example.items():
                        self.static_in[k].copy_(v)
                for _ in range(max(1, warmup)):          # weight images, allocator blocks, side stream: all warm
                    net(dict(self.static_in))
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            n0 = lib.pn2_launch_count()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = net(dict(self.static_in))
            self.launches_per_replay = int(lib.pn2_launch_count() - n0)   # kernels of this library inside the graph

    def __call__(self, pc=None):
        """Enqueue one step on ``self.stream`` (after copying ``pc`` into the static input when given)."""
        with torch.cuda.stream(self.stream):
            if pc is not None:
                for k, v in _as_inputs(pc).items():
                    if v.data_ptr() != self.static_in[k].data_ptr():
                        self.static_in[k].copy_(v, non_blocking=True)
            self.graph.replay()
        return self.out


class BackbonePipeline:
    """Several batches in flight from HOST buffers: the serving loop `bench.py` measures as ``e2e``.

        pipe = BackbonePipeline(net, example_pc, lanes=8)       # one captured step + pinned result buffers per lane
        for pc_host in loader:                                  # (B, N, 3+C) f32, ideally pinned
            ticket = pipe.submit(pc_host)                       # H2D copy, graph replay, D2H of the results: all async
            ...
            out = pipe.result(ticket)                           # dict of pinned host tensors (waits for that batch only)

    ``example`` / ``submit`` take either the reference's (B, N, 3+C) f32 ``point_clouds`` or the compact pair
    ``net.pack_point_clouds(point_clouds)`` = (xyz f32, feature rows bf16): the step is bound by the host-to-device
    copy, the bf16 arm rounds the features to bf16 first thing anyway, so shipping them as bf16 halves the bytes and
    gives bit-identical results (tests/test_fused_gpu.py::test_compact_input_is_bit_identical).

    A lane is reused round robin: ``submit`` first waits for the lane's previous batch, so at most ``lanes`` batches
    are in flight and a ticket's host buffers stay valid until ``lanes`` further submissions.
    """

    def __init__(self, net, example, lanes=8, outputs=("fp2_features", "fp2_xyz", "fp2_inds"), streams=None,
                 sampling_mode=None):
        example = _as_inputs(example)
        first = next(iter(example.values()))
        dev = first.device if first.is_cuda else torch.device("cuda", torch.cuda.current_device())
        example = {k: v.to(dev) for k, v in example.items()}
        self.streams = list(streams) if streams is not None else [lane_stream(dev) for _ in range(lanes)]
        self.inputs = [{k: v.clone() for k, v in example.items()} for _ in self.streams]
        # From host buffers the step is bound by the PCIe copy, not by SM time: the fastest single step ("latency") keeps
        # the lanes short; measured 4.6 k scenes/s end to end against 3.5 k with the throughput-mode sampling kernel.
        if sampling_mode is None:
            sampling_mode = "latency"
        self.sampling_mode = sampling_mode
        self.steps = [GraphedBackbone(net, example, stream=st, static_input=buf, sampling_mode=sampling_mode)
                      for st, buf in zip(self.streams, self.inputs)]
        self.outputs = tuple(outputs)
        self.host = [{k: torch.empty_like(s.out[k], device="cpu").pin_memory() for k in self.outputs} for s in self.steps]
        self.done = [torch.cuda.Event() for _ in self.streams]
        self._busy = [False] * len(self.streams)
        self._next = 0
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in example.values())
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.host[0].values())

    def submit(self, pc):
        ln = self._next
        self._next = (ln + 1) % len(self.streams)
        if self._busy[ln]:
            self.done[ln].synchronize()
        with torch.cuda.stream(self.streams[ln]):
            for k, v in _as_inputs(pc).items():
                self.inputs[ln][k].copy_(v, non_blocking=True)
            out = self.steps[ln]()
            for k, v in self.host[ln].items():
                v.copy_(out[k], non_blocking=True)
            self.done[ln].record()
        self._busy[ln] = True
        return ln

    def result(self, ticket):
        self.done[ticket].synchronize()
        self._busy[ticket] = False
        return self.host[ticket]

    def drain(self):
        for ln, busy in enumerate(self._busy):
            if busy:
                self.done[ln].synchronize()
                self._busy[ln] = False
