"""Correspondences + back-projection (csrc/projection.cu) for all views of a scene, against the reference's per-view
PyTorch sequence run on the same GPU (GPU box).

One JSON line: device time of compute_projection_views and project_views (CUDA events, L2 flushed between
iterations), algorithmic HBM bytes against MEASURED_PEAKS.json, and the per-view loop of lib/projection.py:191-279
restated with the same torch calls (frustum test, mm, rounding, masks with their `.any()` synchronisations,
index_select, index assignment) timed by wall clock, because it is host-synchronous.
"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from situation3d_b200.projection import ProjectionHelper
from situation3d_b200.synthetic import make_scene, make_views

N, V, C = int(os.environ.get("PROJ_N", 50000)), int(os.environ.get("PROJ_V", 64)), 128
pts = make_scene(3, n_points=N, n_features=1)[:, :3].astype(np.float32)
intrinsic, poses, depths = make_views(pts, V, seed=1)
dims, dmin, dmax, acc = [41, 32], 0.4, 4.0, 0.05
helper = ProjectionHelper(torch.from_numpy(intrinsic), dmin, dmax, dims, acc)
points, c2w, depth = torch.from_numpy(pts).cuda(), torch.from_numpy(poses).cuda(), torch.from_numpy(depths).cuda()
w2c = torch.inverse(c2w)
label = torch.randn(V, C, dims[1], dims[0], device="cuda")
flush = torch.empty(1 << 28, dtype=torch.float32, device="cuda")          # 1 GiB > 126 MB of L2


def timed(fn, steps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


i3, i2, counts = helper.compute_projection_views(points, depth, c2w, w2c)
t_cp = timed(lambda: helper.compute_projection_views(points, depth, c2w, w2c))
out_buf = helper.project_views(label, i3, i2, N)
t_pr = timed(lambda: helper.project_views(label, i3, i2, N), steps=5)
del out_buf


# the reference's sequence for one view (lib/projection.py:191-254 and :257-279), on the GPU
intr_t = torch.from_numpy(intrinsic).cuda()
corner_points = helper.corner_points


def ref_view(v):
    cam2world = c2w[v]
    world2cam = torch.inverse(cam2world)
    ind_points = torch.arange(0, N, device="cuda")
    coords = cam2world.new(4, N)
    coords[:3, :] = torch.t(points)
    coords[3, :].fill_(1)
    cc = torch.bmm(cam2world.repeat(8, 1, 1), corner_points.unsqueeze(2))
    normals = cc.new(6, 3)
    for k, (o, a, b) in enumerate([(0, 3, 1), (1, 2, 5), (2, 3, 6), (3, 0, 7), (0, 1, 4), (5, 6, 4)]):
        normals[k] = torch.linalg.cross((cc[a][:3] - cc[o][:3]).view(-1), (cc[b][:3] - cc[o][:3]).view(-1))
    p1, p2 = points - cc[2][:3].view(-1), points - cc[4][:3].view(-1)
    mask = torch.ones(N, device="cuda") > 0
    for k in range(6):
        mask = mask * (torch.round(torch.mm(p1 if k < 3 else p2, normals[k].unsqueeze(1)) * 100) / 100 < 0).squeeze()
    if not mask.any():
        return None
    ind_points = ind_points[mask]
    coords = coords[:, ind_points]
    camera = torch.mm(world2cam, coords)
    camera[0] = (camera[0] * intr_t[0][0]) / camera[2] + intr_t[0][2]
    camera[1] = (camera[1] * intr_t[1][1]) / camera[2] + intr_t[1][2]
    image = torch.round(camera).long()
    valid = torch.ge(image[0], 0) * torch.ge(image[1], 0) * torch.lt(image[0], dims[0]) * torch.lt(image[1], dims[1])
    if not valid.any():
        return None
    vi = image[1][valid] * dims[0] + image[0][valid]
    dv = torch.index_select(depth[v].view(-1), 0, vi)
    dm = dv.ge(dmin) * dv.le(dmax) * torch.abs(dv - camera[2][valid]).le(acc)
    if not dm.any():
        return None
    upd = ind_points[valid][dm]
    r3, r2 = upd.new(N + 1).fill_(0), upd.new(N + 1).fill_(0)
    r3[0] = upd.shape[0]; r2[0] = upd.shape[0]
    r3[1:1 + r3[0]] = upd
    r2[1:1 + r2[0]] = torch.index_select(vi, 0, torch.nonzero(dm)[:, 0])
    return r3, r2


def ref_project(v, r3, r2):
    output = label.new(C, N).fill_(0)
    num_ind = r3[0]
    if num_ind > 0:
        vals = torch.index_select(label[v].view(C, -1), 1, r2[1:1 + num_ind])
        output.view(C, -1)[:, r3[1:1 + num_ind]] = vals
    return output


same, checked = 0, 0
for v in range(V):
    r = ref_view(v)
    checked += 1
    same += int((r is None and int(counts[v]) == 0) or (r is not None and torch.equal(r[0], i3[v]) and torch.equal(r[1], i2[v])))
torch.cuda.synchronize(); t0 = time.perf_counter()
refs = [ref_view(v) for v in range(V)]
torch.cuda.synchronize(); t_ref_cp = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
for v in range(V):
    if refs[v] is not None:
        o = ref_project(v, *refs[v])
torch.cuda.synchronize(); t_ref_pr = (time.perf_counter() - t0) * 1e3
# across-view max-pooling (the multiview feature of every point): one pass here, against the dense project_views + max
# of this library and against the reference's per-frame project() folded with a running torch.max
pooled = helper.project_views_maxpool(label, i3, i2, N)
t_pool = timed(lambda: helper.project_views_maxpool(label, i3, i2, N), steps=5)
dense = helper.project_views(label, i3, i2, N)
pool_same = bool(torch.equal(pooled, torch.clamp_min(dense.max(dim=0)[0], 0.0).t().contiguous()))
del dense
torch.cuda.synchronize(); t0 = time.perf_counter()
acc_ref = label.new_zeros(C, N)
for v in range(V):
    if refs[v] is not None:
        acc_ref = torch.maximum(acc_ref, ref_project(v, *refs[v]))
torch.cuda.synchronize(); t_ref_pool = (time.perf_counter() - t0) * 1e3
pool_same_ref = bool(torch.equal(pooled, acc_ref.t().contiguous()))
b_pool = N * C * 4 + V * (N + 1) * 16 + V * C * dims[0] * dims[1] * 4
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6552.6))
b_cp = V * (N + 1) * 16 + N * 12 + V * dims[0] * dims[1] * 4
b_pr = V * C * N * 4 + V * (N + 1) * 16 + V * C * dims[0] * dims[1] * 4
print(json.dumps({"workload": "projection N=%d views=%d C=%d" % (N, V, C), "correspondences": int(counts.sum()),
                  "compute_projection_ms": round(t_cp, 4), "compute_projection_GBps": round(b_cp / t_cp / 1e6, 1),
                  "compute_projection_frac": round(b_cp / t_cp / 1e6 / peak, 3), "view_points_per_s": round(V * N / t_cp * 1e3),
                  "project_ms": round(t_pr, 4), "project_GBps": round(b_pr / t_pr / 1e6, 1), "project_frac": round(b_pr / t_pr / 1e6 / peak, 3),
                  "maxpool_ms": round(t_pool, 4), "maxpool_GBps": round(b_pool / t_pool / 1e6, 1),
                  "maxpool_equals_dense_project_then_max": pool_same, "maxpool_equals_reference_sequence": pool_same_ref,
                  "reference_sequence_project_and_max_ms": round(t_ref_pool, 2),
                  "hbm_peak_GBps": peak, "reference_sequence_compute_projection_ms": round(t_ref_cp, 2),
                  "reference_sequence_project_ms": round(t_ref_pr, 2),
                  "views_identical_to_reference_sequence": "%d/%d" % (same, checked)}))
