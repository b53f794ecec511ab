"""oracle/voxelize.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the dataset-side hard voxelizer and MeanVFE:
  * reference call site: pcdet/datasets/processor/data_processor.py L115-143
    (`transform_points_to_voxels`) -> `VoxelGeneratorWrapper.generate` L44-60
    -> third-party `spconv.utils.Point2VoxelCPU3d.point_to_voxel` (spconv 2.x,
    un-vendored, un-pinned: setup.py L48, docker/Dockerfile L55).
  * MeanVFE: pcdet/models/backbones_3d/vfe/mean_vfe.py L25-29.

PARITY UNPINNED by the reference (no tests, no golden vectors).  The loop is
restated from spconv's published behaviour; the C
version (voxelize.c) is pinned against the literal Python loop below and
against golden vectors produced through the reference's own
`data_processor.py` (tests/golden/make_golden.py).

May be imported only by tests/, __graft_entry__.smoke() and the CPU-baseline
legs of bench.py.
"""
import ctypes
import math

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        lib.oracle_points_to_voxels.restype = ctypes.c_int
        lib.oracle_points_to_voxels.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_mean_vfe.restype = None
        lib.oracle_mean_vfe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        _lib = lib
    return _lib


def grid_size_xyz(point_cloud_range, voxel_size):
    """data_processor.py L117-118: round((hi - lo) / vsize) as int64, x,y,z order."""
    pcr = np.asarray(point_cloud_range)
    g = (pcr[3:6] - pcr[0:3]) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def points_to_voxels_py(points, vsize_xyz, range_xyz, max_points, max_voxels, legacy_break=False):
    """Literal sequential loop (Appendix C.1).  Slow: small inputs only."""
    points = np.ascontiguousarray(points, dtype=np.float32)
    vs = np.asarray(vsize_xyz, dtype=np.float32)
    rg = np.asarray(range_xyz, dtype=np.float32)
    grid = grid_size_xyz(range_xyz, vsize_xyz)
    n, f = points.shape
    voxels = np.zeros((max_voxels, max_points, f), np.float32)
    coords = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    table = {}
    nv = 0
    for i in range(n):
        c = [0, 0, 0]
        ok = True
        for j in range(3):
            q = np.float32(np.float32(points[i, j] - rg[j]) / vs[j])
            fl = math.floor(float(q)) if np.isfinite(q) else -1
            if fl < 0 or fl >= int(grid[j]):
                ok = False
                break
            c[2 - j] = int(fl)
        if not ok:
            continue
        key = tuple(c)
        v = table.get(key, -1)
        if v == -1:
            if nv >= max_voxels:
                if legacy_break:
                    break
                continue
            v = nv
            nv += 1
            table[key] = v
            coords[v] = c
        if num[v] < max_points:
            voxels[v, num[v]] = points[i]
            num[v] += 1
    return voxels[:nv].copy(), coords[:nv].copy(), num[:nv].copy()


class _Arr:
    """Minimal stand-in for a cumm.tensorview tensor: `.numpy()` returns a copy."""

    def __init__(self, a):
        self._a = a

    def numpy(self):
        return np.array(self._a, copy=True)

    def numpy_view(self):
        return self._a


def from_numpy(a):
    """cumm.tensorview.from_numpy stand-in (data_processor.py L10, L54)."""
    return _Arr(np.ascontiguousarray(a))


class Point2VoxelCPU3d:
    """Same constructor keywords and method as the class data_processor.py L36-42, L54 uses."""

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features,
                 max_num_points_per_voxel, max_num_voxels, legacy_break=False):
        self.vsize = np.asarray(vsize_xyz, dtype=np.float32).copy()
        self.range = np.asarray(coors_range_xyz, dtype=np.float32).copy()
        self.grid = grid_size_xyz(np.asarray(coors_range_xyz), vsize_xyz).astype(np.int32)
        self.f = int(num_point_features)
        self.k = int(max_num_points_per_voxel)
        self.max_voxels = int(max_num_voxels)
        self.legacy_break = bool(legacy_break)
        self._table = None

    def point_to_voxel(self, pc):
        pts = pc.numpy_view() if isinstance(pc, _Arr) else np.asarray(pc)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        assert pts.ndim == 2 and pts.shape[1] == self.f, (pts.shape, self.f)
        lib = _load()
        if self._table is None:
            self._table = np.full(int(np.prod(self.grid.astype(np.int64))), -1, np.int32)
        voxels = np.zeros((self.max_voxels, self.k, self.f), np.float32)
        coords = np.zeros((self.max_voxels, 3), np.int32)
        num = np.zeros((self.max_voxels,), np.int32)
        nv = lib.oracle_points_to_voxels(
            pts.ctypes.data, pts.shape[0], self.f, self.vsize.ctypes.data, self.range.ctypes.data,
            self.grid.ctypes.data, self.k, self.max_voxels, int(self.legacy_break),
            self._table.ctypes.data, voxels.ctypes.data, coords.ctypes.data, num.ctypes.data)
        return _Arr(voxels[:nv]), _Arr(coords[:nv]), _Arr(num[:nv])


def mean_vfe(voxels, num):
    """mean_vfe.py L25-29 on numpy arrays."""
    voxels = np.ascontiguousarray(voxels, np.float32)
    num = np.ascontiguousarray(num, np.int32)
    v, k, f = voxels.shape
    out = np.empty((v, f), np.float32)
    _load().oracle_mean_vfe(voxels.ctypes.data, num.ctypes.data, v, k, f, out.ctypes.data)
    return out


def canonical_voxels(voxels, coords, num, batch_idx=None):
    """Canonical sort used by the parity tests: rows ordered by (b,)z,y,x."""
    coords = np.asarray(coords)
    keys = [coords[:, i] for i in range(coords.shape[1] - 1, -1, -1)]
    if batch_idx is not None:
        keys.append(np.asarray(batch_idx))
    order = np.lexsort(keys)
    return voxels[order], coords[order], num[order], order
