"""Point <-> pixel correspondences and back-projection of image features onto points.

Mirrors ``ProjectionHelper`` of the reference (lib/projection.py:5-279), the geometry half of the multiview-feature
producer (ENet feature maps -> 128-d per point, CONF.MULTIVIEW at lib/config.py:36): same constructor, same method
names, same return formats -- ``compute_projection`` gives the two ``(num_points + 1,)`` int64 lists with the count in
element 0 (or ``None``), ``project`` the zero-filled ``(C, num_points)`` feature array.  The reference spends ~25
PyTorch launches and three host synchronisations per camera view; ``compute_projection_views`` / ``project_views``
take all views of a scene at once (two launches / one dense pass, csrc/projection.cu).

The summation order of the reference's small matrix products is the BLAS's; here it is fixed (see projection.cu), so
the index lists can differ from the reference's only for a point within an ulp of a rounding boundary.  There is no
CPU path: CPU tensors raise.
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


def _host_floats(vals):
    return torch.tensor([float(v) for v in vals], dtype=torch.float32)


@torch.no_grad()
def project_views(label, lin_indices_3d, lin_indices_2d, num_points):
    """label (V,C,H,W) or (V,C,HW); index lists (V,num_points+1) -> (V,C,num_points)."""
    if not (label.is_cuda and lin_indices_3d.is_cuda and lin_indices_2d.is_cuda):
        raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
    V, C = label.shape[0], label.shape[1]
    label = label.to(torch.float32).reshape(V, C, -1).contiguous()
    i3 = lin_indices_3d.to(torch.int64).reshape(V, -1).contiguous()
    i2 = lin_indices_2d.to(torch.int64).reshape(V, -1).contiguous()
    if i3.shape[1] != num_points + 1 or i2.shape[1] != num_points + 1:
        raise RuntimeError("ProjectionHelper.project: index lists must have num_points + 1 entries")
    dev = label.device
    out = torch.empty((V, C, num_points), dtype=torch.float32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.pn2_project_workspace_bytes(V, num_points)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.pn2_project(V, C, label.shape[2], num_points, ptr(label), ptr(i3), ptr(i2), ptr(out), ptr(status), ptr(ws),
                              nbytes, stream_ptr()), "project")
    if int(status.item()) != 0:
        raise IndexError("ProjectionHelper.project: index out of range")
    return out


@torch.no_grad()
def project_views_maxpool(label, lin_indices_3d, lin_indices_2d, num_points, layout="rows"):
    """The multiview feature of every point: max over the views that see it of the back-projected image feature, 0 where
    no view does -- ``project`` of every frame (lib/projection.py:257-279) folded with the running element-wise max into a
    zero-initialised per-point array (the "enet_feats_maxpool" input, lib/config.py:36) in one pass, without the
    (V, C, num_points) intermediate.  label (V,C,H,W) or (V,C,HW); lists (V,num_points+1).
    Returns (num_points, C) for ``layout="rows"`` (the channel-last rows the backbone reads) or (C, num_points)."""
    if not (label.is_cuda and lin_indices_3d.is_cuda and lin_indices_2d.is_cuda):
        raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
    if layout not in ("rows", "channels"):
        raise ValueError("layout must be 'rows' or 'channels'")
    V, C = label.shape[0], label.shape[1]
    label = label.to(torch.float32).reshape(V, C, -1).contiguous()
    i3 = lin_indices_3d.to(torch.int64).reshape(V, -1).contiguous()
    i2 = lin_indices_2d.to(torch.int64).reshape(V, -1).contiguous()
    if V > 0 and (i3.shape[1] != num_points + 1 or i2.shape[1] != num_points + 1):
        raise RuntimeError("ProjectionHelper.project: index lists must have num_points + 1 entries")
    dev = label.device
    out = torch.empty((num_points, C) if layout == "rows" else (C, num_points), dtype=torch.float32, device=dev)
    status = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.pn2_project_workspace_bytes(V, num_points)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.pn2_project_maxpool(V, C, label.shape[2], num_points, ptr(label), ptr(i3), ptr(i2), ptr(out),
                                      1 if layout == "rows" else 0, ptr(status), ptr(ws), nbytes, stream_ptr()), "project_maxpool")
    if int(status.item()) != 0:
        raise IndexError("ProjectionHelper.project: index out of range")
    return out


class ProjectionHelper:
    def __init__(self, intrinsic, depth_min, depth_max, image_dims, accuracy, cuda=True):
        if not cuda:
            raise RuntimeError("ProjectionHelper: cuda=False is not supported (there is no CPU path)")
        self.intrinsic = intrinsic
        self.depth_min = depth_min
        self.depth_max = depth_max
        self.image_dims = image_dims
        self.accuracy = accuracy
        self.cuda = cuda
        self._compute_corner_points()

    # lib/projection.py:17-21
    def depth_to_skeleton(self, ux, uy, depth):
        x = (ux - self.intrinsic[0][2]) / self.intrinsic[0][0]
        y = (uy - self.intrinsic[1][2]) / self.intrinsic[1][1]
        return torch.Tensor([depth * x, depth * y, depth])

    # lib/projection.py:23-26
    def skeleton_to_depth(self, p):
        x = (p[0] * self.intrinsic[0][0]) / p[2] + self.intrinsic[0][2]
        y = (p[1] * self.intrinsic[1][1]) / p[2] + self.intrinsic[1][2]
        return torch.Tensor([x, y, p[2]])

    # lib/projection.py:29-46: the image corners at depth_min, then at depth_max, camera frame, homogeneous
    def _compute_corner_points(self):
        w, h = self.image_dims[0] - 1, self.image_dims[1] - 1
        host = torch.ones(8, 4)
        for k, (ux, uy, d) in enumerate([(0, 0, self.depth_min), (w, 0, self.depth_min), (w, h, self.depth_min), (0, h, self.depth_min),
                                         (0, 0, self.depth_max), (w, 0, self.depth_max), (w, h, self.depth_max), (0, h, self.depth_max)]):
            host[k, :3] = self.depth_to_skeleton(ux, uy, d)
        self._corner_host = host[:, :3].contiguous()
        self._intr4 = _host_floats([self.intrinsic[0][0], self.intrinsic[1][1], self.intrinsic[0][2], self.intrinsic[1][2]])
        self._range3 = _host_floats([self.depth_min, self.depth_max, self.accuracy])
        self.corner_points = host.cuda()

    def _poses(self, camera_to_world):
        if not camera_to_world.is_cuda:
            raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
        c2w = camera_to_world.to(torch.float32).reshape(-1, 4, 4).contiguous()
        return c2w

    def _planes(self, camera_to_world, want_corners, want_normals):
        c2w = self._poses(camera_to_world)
        V = c2w.shape[0]
        corners = torch.empty((V, 8, 4), dtype=torch.float32, device=c2w.device) if want_corners else None
        normals = torch.empty((V, 6, 3), dtype=torch.float32, device=c2w.device) if want_normals else None
        with torch.cuda.device(c2w.device):
            check(lib.pn2_frustum_planes(V, ptr(c2w), ptr(self._intr4), ptr(self._range3), int(self.image_dims[0]), int(self.image_dims[1]),
                                         ptr(self._corner_host), ptr(corners) if want_corners else None,
                                         ptr(normals) if want_normals else None, stream_ptr()), "frustum_planes")
        return corners, normals

    def compute_frustum_corners(self, camera_to_world):
        """(4,4) pose -> (8,4,1) world coordinates of the frustum corners (lib/projection.py:48-70)."""
        return self._planes(camera_to_world, True, False)[0][0].unsqueeze(2)

    def compute_frustum_normals(self, corner_coords):
        """(8,4[,1]) corners -> (6,3) inward plane normals (lib/projection.py:72-119), by the device routine that
        compute_projection uses (the corners go in as corner points under the identity pose, which is exact)."""
        if not corner_coords.is_cuda:
            raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
        host = corner_coords.reshape(8, 4)[:, :3].to(torch.float32).cpu().contiguous()
        eye = torch.eye(4, dtype=torch.float32, device=corner_coords.device).reshape(1, 4, 4)
        normals = torch.empty((1, 6, 3), dtype=torch.float32, device=corner_coords.device)
        with torch.cuda.device(corner_coords.device):
            check(lib.pn2_frustum_planes(1, ptr(eye), ptr(self._intr4), ptr(self._range3), int(self.image_dims[0]),
                                         int(self.image_dims[1]), ptr(host), None, ptr(normals), stream_ptr()), "frustum_planes")
        return normals[0]

    def points_in_frustum(self, corner_coords, normals, new_pts, return_mask=False):
        """Boolean mask (``return_mask=True``) or number of ``new_pts`` (M,3) inside the frustum given by its corners
        (8,4[,1]) and plane normals (6,3) (lib/projection.py:121-155)."""
        if not (corner_coords.is_cuda and normals.is_cuda):
            raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
        dev = corner_coords.device
        pts = new_pts.to(dev, torch.float32).contiguous()            # the reference moves new_pts to the GPU itself (:133)
        cc = corner_coords.reshape(8, 4).to(torch.float32).contiguous()
        nr = normals.reshape(6, 3).to(torch.float32).contiguous()
        m = pts.shape[0]
        mask = torch.empty(m, dtype=torch.uint8, device=dev) if return_mask else None
        count = torch.empty(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.pn2_points_in_frustum(m, ptr(pts), ptr(cc), ptr(nr), ptr(mask), ptr(count), stream_ptr()), "points_in_frustum")
        return mask.bool() if return_mask else count[0].long()

    def compute_projection_views(self, points, depth, camera_to_world, world_to_camera=None):
        """All views of a scene at once.  points (N,3), depth (V,H,W), camera_to_world (V,4,4), all CUDA f32.
        Returns (indices_3d (V,N+1) int64, indices_2d (V,N+1) int64, counts (V,) int32) -- row v is what the
        reference's compute_projection returns for view v, a count of 0 standing for its ``None``."""
        c2w = self._poses(camera_to_world)
        if not (points.is_cuda and depth.is_cuda):
            raise RuntimeError("ProjectionHelper: CUDA tensors required (there is no CPU path)")
        V, N = c2w.shape[0], points.shape[0]
        W, H = int(self.image_dims[0]), int(self.image_dims[1])
        points = points.to(torch.float32).contiguous()
        depth = depth.to(torch.float32).reshape(V, -1).contiguous()
        if points.dim() != 2 or points.shape[1] != 3 or depth.shape[1] != W * H:
            raise RuntimeError("ProjectionHelper: expected points (N,3) and depth (V,%d,%d)" % (H, W))
        w2c = torch.inverse(c2w) if world_to_camera is None else world_to_camera.to(torch.float32).reshape(V, 4, 4)               # :203
        w2c = w2c.contiguous()          # torch.inverse hands back column-major batches
        dev = points.device
        i3 = torch.empty((V, N + 1), dtype=torch.int64, device=dev)
        i2 = torch.empty((V, N + 1), dtype=torch.int64, device=dev)
        counts = torch.empty(V, dtype=torch.int32, device=dev)
        nbytes = lib.pn2_compute_projection_workspace_bytes(V, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(lib.pn2_compute_projection(V, N, ptr(points), ptr(depth), ptr(c2w), ptr(w2c), ptr(self._intr4), ptr(self._range3),
                                             W, H, ptr(self._corner_host), ptr(i3), ptr(i2), ptr(counts), ptr(ws), nbytes,
                                             stream_ptr()), "compute_projection")
        return i3, i2, counts

    def compute_projection(self, points, depth, camera_to_world):
        """One view, the reference's signature and return value (lib/projection.py:191-254): (indices_3d, indices_2d)
        or None when no point corresponds to a pixel."""
        i3, i2, counts = self.compute_projection_views(points, depth.reshape(1, -1), camera_to_world.reshape(1, 4, 4))
        if int(counts.item()) == 0:                       # the reference synchronises three times here (`mask.any()`)
            return None
        return i3[0], i2[0]

    def project_views(self, label, lin_indices_3d, lin_indices_2d, num_points):
        """label (V,C,H,W) or (V,C,HW); index lists (V,num_points+1) -> (V,C,num_points)."""
        return project_views(label, lin_indices_3d, lin_indices_2d, num_points)

    @torch.no_grad()
    def project_views_maxpool(self, label, lin_indices_3d, lin_indices_2d, num_points, layout="rows"):
        """All views -> the max-pooled multiview feature of every point (see the module-level function)."""
        return project_views_maxpool(label, lin_indices_3d, lin_indices_2d, num_points, layout)

    def project(self, label, lin_indices_3d, lin_indices_2d, num_points):
        """One view, the reference's signature (lib/projection.py:257-279): label (C,H,W) or (H,W) -> (C,num_points)."""
        c = 1 if label.dim() == 2 else label.shape[0]
        return self.project_views(label.reshape(1, c, -1), lin_indices_3d.reshape(1, -1), lin_indices_2d.reshape(1, -1), num_points)[0]


class Projection(torch.autograd.Function):
    """The reference's autograd wrapper of the back-projection (lib/projection.py:283-309).  forward is
    ProjectionHelper.project.  The reference's backward cannot run (``save_for_backward`` is commented out at :296,
    so ``ctx.saved_variables`` is empty, and the gradient shape is hard-wired to 32 x 41); it is not reproduced."""

    @staticmethod
    def forward(ctx, label, lin_indices_3d, lin_indices_2d, num_points):
        c = 1 if label.dim() == 2 else label.shape[0]
        return project_views(label.reshape(1, c, -1), lin_indices_3d.reshape(1, -1), lin_indices_2d.reshape(1, -1), num_points)[0]

    @staticmethod
    def backward(ctx, grad_output):
        raise NotImplementedError("Projection.backward: the reference's backward is unreachable (lib/projection.py:296)")
