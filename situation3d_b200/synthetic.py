"""Synthetic ScanNet-shaped scenes: there is no dataset access, so benches and
parity tests use seeded rooms -- floor + walls + box-shaped furniture, metre-scale, centred on
the origin, randomly permuted, with exact duplicates and all-zero padding points so that the
FPS skip rule and every tie-break are exercised.  CPU/numpy only; deterministic per seed."""
import numpy as np


def make_scene(seed, n_points=40000, n_features=129, dtype=np.float32):
    """One scene: (n_points, 3 + n_features) float32 = xyz | height | relu(N(0,1)) features."""
    rng = np.random.default_rng(1000 + seed)
    lx, ly, h = rng.uniform(4, 9), rng.uniform(3, 7), 2.6
    n_struct = int(0.6 * n_points)
    n_furn = n_points - n_struct

    # floor and four walls, area-weighted
    areas = np.array([lx * ly, lx * h, lx * h, ly * h, ly * h])
    which = rng.choice(5, size=n_struct, p=areas / areas.sum())
    u, v = rng.uniform(0, 1, n_struct), rng.uniform(0, 1, n_struct)
    pts = np.zeros((n_struct, 3))
    pts[which == 0] = np.stack([u * lx, v * ly, 0 * u], 1)[which == 0]
    pts[which == 1] = np.stack([u * lx, 0 * u, v * h], 1)[which == 1]
    pts[which == 2] = np.stack([u * lx, 0 * u + ly, v * h], 1)[which == 2]
    pts[which == 3] = np.stack([0 * u, u * ly, v * h], 1)[which == 3]
    pts[which == 4] = np.stack([0 * u + lx, u * ly, v * h], 1)[which == 4]

    # furniture: points on the surfaces of axis-aligned boxes standing on the floor
    n_box = int(rng.integers(12, 26))
    sizes = rng.uniform(0.3, 2.0, (n_box, 3))
    sizes[:, 2] = np.minimum(sizes[:, 2], h - 0.1)
    origin = np.stack([rng.uniform(0, np.maximum(lx - sizes[:, 0], 0.1)),
                       rng.uniform(0, np.maximum(ly - sizes[:, 1], 0.1)), np.zeros(n_box)], 1)
    surf = 2 * (sizes[:, 0] * sizes[:, 1] + sizes[:, 0] * sizes[:, 2] + sizes[:, 1] * sizes[:, 2])
    box = rng.choice(n_box, size=n_furn, p=surf / surf.sum())
    q = rng.uniform(0, 1, (n_furn, 3))
    face = rng.integers(0, 6, n_furn)
    axis, side = face // 2, face % 2
    q[np.arange(n_furn), axis] = side
    fpts = origin[box] + q * sizes[box]

    xyz = np.concatenate([pts, fpts], 0)
    xyz += rng.normal(0, 0.005, xyz.shape)
    xyz -= np.array([lx / 2, ly / 2, 0.0])          # room centre at the origin (ScanNet-aligned)
    xyz = xyz[rng.permutation(n_points)]
    # exact duplicates (0.5 %) and 8 all-zero padding points
    n_dup = max(1, n_points // 200)
    dst = rng.choice(n_points, n_dup, replace=False)
    xyz[dst] = xyz[rng.choice(n_points, n_dup, replace=False)]
    xyz[rng.choice(n_points, min(8, n_points), replace=False)] = 0.0

    out = np.empty((n_points, 3 + n_features), dtype=dtype)
    out[:, :3] = xyz
    if n_features > 0:
        out[:, 3] = xyz[:, 2] - np.quantile(xyz[:, 2], 0.01)          # height above the floor
        if n_features > 1:
            out[:, 4:] = np.maximum(rng.normal(0, 1, (n_points, n_features - 1)), 0)   # post-ReLU multiview features
    return out


def make_batch(batch, n_points=40000, n_features=129, first_seed=0):
    """(batch, n_points, 3 + n_features) float32."""
    return np.stack([make_scene(first_seed + s, n_points, n_features) for s in range(batch)], 0)


def make_situations(batch, seed=0):
    """(batch, 7) float32: translation inside a room, unit quaternion (xyzw) of a z-rotation."""
    rng = np.random.default_rng(5000 + seed)
    t = np.stack([rng.uniform(-2, 2, batch), rng.uniform(-1.5, 1.5, batch), np.zeros(batch)], 1)
    a = rng.uniform(-np.pi, np.pi, batch)
    q = np.stack([np.zeros(batch), np.zeros(batch), np.sin(a / 2), np.cos(a / 2)], 1)
    return np.concatenate([t, q], 1).astype(np.float32)


def randomize_bn_stats(module, seed=0):
    """Give every BatchNorm non-trivial running statistics and affine parameters so that
    eval-mode folding is exercised (module default init has mean 0 / var 1 / gamma 1 / beta 0)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
    return module


# ---- camera views of a scene, for the correspondence / back-projection path (projection.py) ---------------
def look_at(eye, target):
    """camera_to_world (4,4) float32 of a pinhole camera (x right, y down, z forward) at `eye` looking at `target`."""
    eye, target = np.asarray(eye, dtype=np.float64), np.asarray(target, dtype=np.float64)
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, [0.0, 0.0, 1.0])
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, d, f, eye
    return m.astype(np.float32)


def make_views(points, n_views, seed=0, image_dims=(41, 32), focal=(37.01983, 38.52470), centre=(20.0, 15.5),
               noise=0.02, holes=0.08):
    """Random camera poses inside the scene's bounding box and, per view, a depth map that is a z-buffer of the
    cloud itself (float64), with Gaussian noise and dropped pixels -- so that every stage of compute_projection
    rejects some points.  Returns (intrinsic (4,4), poses (V,4,4), depths (V,H,W)), float32 numpy."""
    rng = np.random.default_rng(seed)
    pts = np.asarray(points, dtype=np.float64)[:, :3]
    lo, hi = pts.min(0), pts.max(0)
    lo, hi = np.where(hi - lo < 1.0, lo - 0.5, lo), np.where(hi - lo < 1.0, hi + 0.5, hi)     # degenerate clouds
    W, H = image_dims
    intrinsic = np.eye(4, dtype=np.float32)
    intrinsic[0, 0], intrinsic[1, 1], intrinsic[0, 2], intrinsic[1, 2] = focal[0], focal[1], centre[0], centre[1]
    hom = np.concatenate([pts, np.ones((pts.shape[0], 1))], 1).T
    poses, depths = [], []
    for _ in range(n_views):
        eye = lo + (hi - lo) * rng.uniform(0.25, 0.75, 3)
        target = lo + (hi - lo) * rng.uniform(0.0, 1.0, 3)
        c2w = look_at(eye, target)
        cam = np.linalg.inv(c2w.astype(np.float64)) @ hom
        with np.errstate(divide="ignore", invalid="ignore"):
            u = np.rint(cam[0] * float(intrinsic[0, 0]) / cam[2] + float(intrinsic[0, 2]))
            v = np.rint(cam[1] * float(intrinsic[1, 1]) / cam[2] + float(intrinsic[1, 2]))
        ok = (cam[2] > 0.05) & (u >= 0) & (v >= 0) & (u < W) & (v < H)
        depth = np.full(W * H, np.inf)
        np.minimum.at(depth, (v[ok] * W + u[ok]).astype(np.int64), cam[2][ok])
        depth[~np.isfinite(depth)] = 0.0
        depth = depth + rng.normal(0.0, noise, depth.shape) * (depth > 0)
        depth[rng.uniform(size=depth.shape) < holes] = 0.0
        poses.append(c2w)
        depths.append(depth.reshape(H, W).astype(np.float32))
    return intrinsic, np.stack(poses), np.stack(depths)
