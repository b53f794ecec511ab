"""`spconv.utils` / `cumm.tensorview`-shaped surface for the dataset-side voxelizer call
(pcdet/datasets/processor/data_processor.py L10, L25, L36-42, L54-59).

`Point2VoxelCPU3d` keeps the reference's class name and constructor keywords so that an unmodified
VoxelGeneratorWrapper binds to it, but the work runs on the GPU: host points -> pinned H2D -> K1 -> D2H.
It must be used from the rank's main process (CUDA in forked DataLoader workers is unsafe); the
throughput path is the in-model GPU voxelizer behind `transform_points_to_voxels_placeholder`
(toda_b200.pcdet_plugin.MeanVFE).
"""
import numpy as np
import torch

from .. import ops


class HostArray:
    """Stand-in for a cumm.tensorview tensor: `.numpy()` copies, `.numpy_view()` aliases."""

    def __init__(self, a):
        self._a = a

    def numpy(self):
        return np.array(self._a, copy=True)

    def numpy_view(self):
        return self._a


def from_numpy(a):
    return HostArray(np.ascontiguousarray(a))


class Point2VoxelCPU3d:
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_points_per_voxel, max_num_voxels,
                 device=None, order=ops.ORDER_FIRST_APPEARANCE):
        self.vsize = [float(v) for v in vsize_xyz]
        self.range = [float(v) for v in coors_range_xyz]
        self.grid = ops.grid_size_xyz(self.range, self.vsize)
        self.f = int(num_point_features)
        self.k = int(max_num_points_per_voxel)
        self.max_voxels = int(max_num_voxels)
        self.order = order
        if not torch.cuda.is_available():
            raise RuntimeError("toda_b200 voxelizer needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self._pin = None

    def _to_device(self, pts):
        n = pts.shape[0]
        if self._pin is None or self._pin.shape[0] < n:
            self._pin = torch.empty((max(n, 1) * 5 // 4, self.f), dtype=torch.float32).pin_memory()
        self._pin[:n].copy_(torch.from_numpy(pts))
        return self._pin[:n].to(self.device, non_blocking=True)

    def point_to_voxel(self, pc):
        pts = pc.numpy_view() if isinstance(pc, HostArray) else np.asarray(pc)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        if pts.ndim != 2 or pts.shape[1] != self.f:
            raise ValueError(f"points must be (N, {self.f}), got {pts.shape}")
        dpts = self._to_device(pts)
        voxels, coords, num, _ = ops.voxelize(dpts, [0, pts.shape[0]], self.range, self.vsize, self.k, self.max_voxels,
                                              num_features=self.f, order=self.order, grid=self.grid)
        return (HostArray(voxels.cpu().numpy()), HostArray(coords[:, 1:].contiguous().cpu().numpy()),
                HostArray(num.cpu().numpy()))

    # spconv 2.x also exposes this spelling on its GPU class
    point_to_voxel_hash = point_to_voxel


Point2VoxelGPU3d = Point2VoxelCPU3d
