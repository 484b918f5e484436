"""Golden vectors of the point <-> pixel correspondences from the REFERENCE's own class (build container only).

    python tests/golden/make_ref_projection_goldens.py        # writes tests/golden/ref_projection.npz

/root/reference/lib/projection.py is loaded unmodified (it imports only torch) and ProjectionHelper is run on the
CPU -- ``Tensor.cuda()`` is the identity while it runs.  A synthetic room (points on the floor, the walls and a few
boxes) is seen from several camera poses; each view's depth map is a z-buffer of the cloud itself with holes and
noise, so that every stage of compute_projection (frustum, image range, depth range, depth agreement) rejects some
points.  Stored: inputs, the reference's index lists, frustum corners / normals / point masks, and project() outputs.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/lib/projection.py"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

IMAGE_DIMS = [41, 32]
DEPTH_MIN, DEPTH_MAX, ACCURACY = 0.4, 4.0, 0.05


def make_intrinsic():
    # a 41 x 32 pinhole camera (the ScanNet colour intrinsics scaled to the feature-map size)
    m = torch.eye(4)
    m[0][0], m[1][1], m[0][2], m[1][2] = 37.01983, 38.52470, 20.0, 15.5
    return m


def room(g, n):
    """Points on the surfaces of a 6 x 5 x 2.6 m room with three boxes in it."""
    parts = []
    per = n // 8
    u = lambda k: torch.rand(k, generator=g)
    parts.append(torch.stack([u(per) * 6 - 3, u(per) * 5 - 2.5, torch.zeros(per)], 1))             # floor
    parts.append(torch.stack([u(per) * 6 - 3, u(per) * 5 - 2.5, torch.full((per,), 2.6)], 1))      # ceiling
    for sx in (-3.0, 3.0):
        parts.append(torch.stack([torch.full((per,), sx), u(per) * 5 - 2.5, u(per) * 2.6], 1))
    for sy in (-2.5, 2.5):
        parts.append(torch.stack([u(per) * 6 - 3, torch.full((per,), sy), u(per) * 2.6], 1))
    rest = n - 6 * per
    c = torch.tensor([[1.0, 0.5, 0.4], [-1.2, -0.8, 0.5], [0.3, -1.5, 0.3]])[torch.randint(0, 3, (rest,), generator=g)]
    parts.append(c + (torch.rand(rest, 3, generator=g) - 0.5) * torch.tensor([0.8, 0.8, 0.8]))
    pts = torch.cat(parts)
    return pts[torch.randperm(pts.shape[0], generator=g)].contiguous()


def look_at(eye, target):
    """camera_to_world of a pinhole camera (x right, y down, z forward) at `eye` looking at `target`."""
    f = target - eye
    f = f / f.norm()
    r = torch.linalg.cross(f, torch.tensor([0.0, 0.0, 1.0]))
    r = r / r.norm()
    d = torch.linalg.cross(f, r)
    m = torch.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, d, f, eye
    return m


def zbuffer(points, c2w, intrinsic, g):
    w2c = torch.inverse(c2w.double())
    cam = (w2c @ torch.cat([points.double(), torch.ones(points.shape[0], 1, dtype=torch.float64)], 1).T)
    u = torch.round(cam[0] * float(intrinsic[0][0]) / cam[2] + float(intrinsic[0][2])).long()
    v = torch.round(cam[1] * float(intrinsic[1][1]) / cam[2] + float(intrinsic[1][2])).long()
    ok = (cam[2] > 0.05) & (u >= 0) & (v >= 0) & (u < IMAGE_DIMS[0]) & (v < IMAGE_DIMS[1])
    depth = torch.full((IMAGE_DIMS[1] * IMAGE_DIMS[0],), 1e9, dtype=torch.float64)
    depth.scatter_reduce_(0, (v * IMAGE_DIMS[0] + u)[ok], cam[2][ok], reduce="amin")
    depth[depth > 1e8] = 0.0                                                     # no surface: invalid depth
    depth = depth + torch.randn(depth.shape, generator=g, dtype=torch.float64) * 0.02 * (depth > 0)
    depth[torch.rand(depth.shape, generator=g) < 0.08] = 0.0                     # sensor holes
    return depth.float().view(IMAGE_DIMS[1], IMAGE_DIMS[0])


def main():
    spec = importlib.util.spec_from_file_location("ref_projection", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from oracle import pn2_oracle as orc

    g = torch.Generator().manual_seed(2024)
    intrinsic = make_intrinsic()
    helper = ref.ProjectionHelper(intrinsic, DEPTH_MIN, DEPTH_MAX, IMAGE_DIMS, ACCURACY)
    out = {"intrinsic": intrinsic.numpy(), "params": np.array([DEPTH_MIN, DEPTH_MAX, ACCURACY], dtype=np.float32),
           "image_dims": np.array(IMAGE_DIMS), "corner_points": helper.corner_points[:, :3].numpy()}
    n = 6000
    points = room(g, n)
    views = 10
    poses, depths, i3s, i2s, corners, normals, masks = [], [], [], [], [], [], []
    for v in range(views):
        eye = torch.tensor([float(torch.rand(1, generator=g) * 4 - 2), float(torch.rand(1, generator=g) * 3 - 1.5),
                            float(torch.rand(1, generator=g) * 1.2 + 0.8)])
        target = torch.tensor([float(torch.rand(1, generator=g) * 6 - 3), float(torch.rand(1, generator=g) * 5 - 2.5),
                               float(torch.rand(1, generator=g) * 1.5)])
        c2w = look_at(eye, target)
        if v == views - 1:
            c2w = look_at(torch.tensor([10.0, 10.0, 1.0]), torch.tensor([20.0, 20.0, 1.0]))     # sees nothing: None
        depth = zbuffer(points, c2w, intrinsic, g)
        res = helper.compute_projection(points, depth, c2w)
        if res is None:
            i3, i2 = torch.zeros(n + 1, dtype=torch.int64), torch.zeros(n + 1, dtype=torch.int64)
        else:
            i3, i2 = res
        cc = helper.compute_frustum_corners(c2w)
        corners.append(cc.squeeze(-1)); normals.append(helper.compute_frustum_normals(cc))
        masks.append(helper.points_in_frustum(cc, normals[-1], points, return_mask=True))
        assert int(helper.points_in_frustum(cc, normals[-1], points)) == int(masks[-1].sum())
        poses.append(c2w); depths.append(depth); i3s.append(i3); i2s.append(i2)
        mine = orc.compute_projection(points, depth, c2w, torch.inverse(c2w), intrinsic, DEPTH_MIN, DEPTH_MAX, IMAGE_DIMS, ACCURACY)
        same = (mine is None and res is None) or (mine is not None and res is not None and torch.equal(mine[0], i3) and torch.equal(mine[1], i2))
        print("view %d: %d correspondences, oracle identical: %s" % (v, int(i3[0]), same))
    label = torch.randn(views, 8, IMAGE_DIMS[1], IMAGE_DIMS[0], generator=g)
    proj = torch.stack([helper.project(label[v], i3s[v], i2s[v], n) for v in range(views)])
    out.update({"points": points.numpy(), "poses": torch.stack(poses).numpy(), "depths": torch.stack(depths).numpy(),
                "indices_3d": torch.stack(i3s).numpy().astype(np.int32), "indices_2d": torch.stack(i2s).numpy().astype(np.int32),
                "corners": torch.stack(corners).numpy(), "normals": torch.stack(normals).numpy(),
                "label": label.numpy(), "frustum_masks": np.packbits(torch.stack(masks).numpy(), axis=1), "project_view3": proj[3].numpy(),
                "project_checksum": proj.double().sum((1, 2)).numpy()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_projection.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
