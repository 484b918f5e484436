"""Voxel-coordinate positional embedding (csrc/voxel_pe.cu) at the reference's shape, B x 5000 x 1408 (GPU box).

Prints one JSON line: device time per call (CUDA events, inputs rotated through buffers larger than L2), the
algorithmic HBM bytes (feature row read + written; coordinates; the 480 KB table is L2-resident) against the
measured copy bandwidth in MEASURED_PEAKS.json, and the reference's own per-sample CPU loop
(blip2_t5.py:107-118 restated with the same torch calls, host tensors in, GPU tensor out) timed beside it.
"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from situation3d_b200.voxel_pe import VoxelPositionalEmbedding

B, P, C = int(os.environ.get("PE_B", 8)), 5000, 1408
m = VoxelPositionalEmbedding().cuda()
g = torch.Generator().manual_seed(0)
nbuf = 4                                                    # 4 x (225 MB in + 225 MB out) >> 126 MB of L2
feats = [torch.randn(B, P, C, generator=g).cuda() for _ in range(nbuf)]
pcs = [torch.randint(0, 256, (B, P, 3), generator=g).float().cuda() for _ in range(nbuf)]
outs = [torch.empty(B, P, C, device="cuda") for _ in range(nbuf)]
from situation3d_b200.voxel_pe import voxel_pe


def run(i, mode="add"):
    return voxel_pe(feats[i % nbuf], pcs[i % nbuf], m.pos_embedding, mode=mode, out=outs[i % nbuf] if mode == "add" else None,
                    validate=False)


res = {}
for mode in ("add", "cat"):
    for i in range(4):
        run(i, mode)
    torch.cuda.synchronize()
    steps = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        run(i, mode)
    e1.record(); torch.cuda.synchronize()
    res[mode] = e0.elapsed_time(e1) / steps
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6552.6))
bytes_add = B * P * (2 * C * 4 + 12)
bytes_cat = B * P * (3 * C * 4 + 12)

# the reference's loop: CPU indexing per sample, host tensor, copy, add on the GPU
pc_host, feat_dev, table_host = pcs[0].cpu(), feats[0], m.pos_embedding.cpu()


def ref_loop():
    pc = pc_host.long()
    all_pcs = torch.zeros((B, P, C))
    for j in range(B):
        all_pcs[j][:, :1407] = torch.cat([table_host[pc[j][:, i]] for i in range(3)], -1)
    return feat_dev + 0.01 * all_pcs.cuda()


ref_loop(); torch.cuda.synchronize(); ts = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); want = ref_loop(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
got, _ = voxel_pe(feats[0], pcs[0], m.pos_embedding, validate=False)
print(json.dumps({"workload": "voxel_pe B=%d P=%d C=%d" % (B, P, C), "add_ms": round(res["add"], 4), "cat_ms": round(res["cat"], 4),
                  "add_GBps": round(bytes_add / res["add"] / 1e6, 1), "cat_GBps": round(bytes_cat / res["cat"] / 1e6, 1),
                  "hbm_peak_GBps": peak, "add_frac": round(bytes_add / res["add"] / 1e6 / peak, 3),
                  "points_per_s": round(B * P / res["add"] * 1e3), "reference_loop_ms": round(1e3 * sorted(ts)[1], 2),
                  "bit_exact_vs_reference_loop": bool(torch.equal(got, want))}))
