"""Device times of the other BASELINE.json configurations on one GPU (GPU box): config 1 (B=1, fp32), config 3
(backbone + situation re-encoding of 256 tokens, B=4 = one GPU's share of 32 scenes over 8 GPUs), config 5
(100k-200k-point scenes, SA1 npoint 4096 / nsample 64, B=8 per GPU)."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from situation3d_b200.backbone_module import Pointnet2Backbone
from situation3d_b200.scene_encoder import SituatedSceneEncoder
from situation3d_b200.synthetic import make_batch, make_situations, randomize_bn_stats

_flush = None


def timeit(fn, reps=5, cushion=False):
    """Median device time (ms).  cushion=True first enqueues a 1 GB memset so that the host has enqueued fn's
    launches before the GPU reaches them (device time of short kernels instead of launch latency) and L2 is cold."""
    global _flush
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        if cushion:
            if _flush is None:
                _flush = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
            _flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))

def graph_time(fn, n=20, reps=5):
    """Device time per call (ms) of a short launch sequence: n calls captured in one CUDA graph and replayed, so that the
    host's enqueue cost (Python + ctypes + allocator, tens of microseconds) is not what the events bracket."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(n):
            fn()
    ts = []
    with torch.cuda.stream(st):
        g.replay(); st.synchronize()
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); g.replay(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e) / n)
    return float(np.median(ts))


out = []
torch.manual_seed(0)
with torch.no_grad():
    # config 1
    net = randomize_bn_stats(Pointnet2Backbone(129, precision="fp32")).eval().cuda()
    pc = torch.from_numpy(make_batch(1, 40000, 129)).cuda()
    ms = timeit(lambda: net({"point_clouds": pc}))
    out.append({"config": 1, "what": "Pointnet2Backbone forward, B=1, 40k points, fp32 arm (FFMA kernels)", "ms": ms, "scenes_per_s": 1e3 / ms})
    net.set_precision("bf16")
    ms = timeit(lambda: net({"point_clouds": pc}))
    out.append({"config": 1, "what": "same scene, bf16 arm", "ms": ms, "scenes_per_s": 1e3 / ms})
    # config 3
    enc = SituatedSceneEncoder(129, 256, precision="bf16").eval().cuda()
    randomize_bn_stats(enc.backbone_net)
    pc4 = torch.from_numpy(make_batch(4, 40000, 129)).cuda()
    sit = torch.from_numpy(make_situations(4)).cuda()
    ms = timeit(lambda: enc({"point_clouds": pc4, "auxiliary_task": sit}))
    out.append({"config": 3, "what": "backbone + re-encoding of 256 tokens, B=4 scenes (one GPU's share of 32 over 8 GPUs), single stream", "ms": ms, "scenes_per_s": 4e3 / ms})
    d = enc({"point_clouds": pc4, "auxiliary_task": sit})
    tok, pos = d["scene_feat"], d["scene_positions"]
    ms = graph_time(lambda: enc.reencoder({"scene_feat": tok, "scene_positions": pos, "auxiliary_task": sit}))
    out.append({"config": 3, "what": "re-encoding alone (transform + pos_embed + prior: 2 launches), 4 x 256 tokens, device time per call (graph replay of 20 calls)", "ms": ms})
    tok32, pos32, sit32 = tok.repeat(8, 1, 1), pos.repeat(8, 1, 1), sit.repeat(8, 1)
    ms = graph_time(lambda: enc.reencoder({"scene_feat": tok32, "scene_positions": pos32, "auxiliary_task": sit32}))
    out.append({"config": 3, "what": "re-encoding alone, 32 x 256 tokens (16.8 MB in+out, 0.54 GFLOP), device time per call (graph replay of 20 calls, L2-warm)",
                "ms": ms, "gbs": 16.8e-3 / (ms * 1e-3), "fp32_tflops": 0.54e-3 / (ms * 1e-3)})
    # config 5
    for n in (100000, 150000, 200000):
        net5 = randomize_bn_stats(Pointnet2Backbone(129, precision="bf16", npoints=(4096, 2048, 1024, 512))).eval().cuda()
        pc5 = torch.from_numpy(make_batch(8, n, 129)).cuda()
        ms = timeit(lambda: net5({"point_clouds": pc5}), reps=3)
        out.append({"config": 5, "what": "stress: B=8 scenes of %d points, npoint 4096/2048/1024/512, nsample 64/32/16/16, bf16, single stream" % n,
                    "ms": ms, "scenes_per_s": 8e3 / ms})
        del pc5, net5
        torch.cuda.empty_cache()
for r in out:
    print(json.dumps(r))
