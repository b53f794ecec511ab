"""Pins the C voxelizer oracle (oracle/voxelize.c) against the literal Python loop and against
hand-derived edge cases: caps, boundary points, max_voxels overflow,
duplicates, empty input."""
import numpy as np
import pytest

from oracle import voxelize as V

VS = [0.5, 0.5, 0.25]
RG = [-8.0, -8.0, -1.5, 8.0, 8.0, 1.5]


def run_c(pts, vs=VS, rg=RG, k=3, mv=500, legacy=False):
    g = V.Point2VoxelCPU3d(vs, rg, pts.shape[1], k, mv, legacy_break=legacy)
    return [t.numpy() for t in g.point_to_voxel(V.from_numpy(pts))]


def rand_pts(seed, n, f=5):
    rng = np.random.default_rng(seed)
    p = rng.random((n, f), dtype=np.float32)
    p[:, :3] = p[:, :3] * np.array([20, 20, 4], np.float32) - np.array([10, 10, 2], np.float32)
    return p


@pytest.mark.parametrize("seed,n,k,mv", [(0, 2000, 3, 500), (1, 3000, 1, 10000), (2, 500, 10, 50), (3, 1, 5, 5)])
def test_c_matches_python_loop(seed, n, k, mv):
    pts = rand_pts(seed, n)
    a = V.points_to_voxels_py(pts, VS, RG, k, mv)
    b = run_c(pts, k=k, mv=mv)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)
    # the dense table must have been restored: a second call gives the same answer
    g = V.Point2VoxelCPU3d(VS, RG, 5, k, mv)
    r1 = [t.numpy() for t in g.point_to_voxel(V.from_numpy(pts))]
    r2 = [t.numpy() for t in g.point_to_voxel(V.from_numpy(pts))]
    for x, y in zip(r1, r2):
        assert np.array_equal(x, y)


def test_legacy_break_differs_only_after_overflow():
    pts = rand_pts(5, 4000)
    a = V.points_to_voxels_py(pts, VS, RG, 3, 100, legacy_break=True)
    b = run_c(pts, k=3, mv=100, legacy=True)
    c = run_c(pts, k=3, mv=100, legacy=False)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert np.array_equal(b[1], c[1])          # same voxels are opened ...
    assert (c[2] >= b[2]).all() and c[2].sum() > b[2].sum()   # ... but 2.x keeps filling them


def test_empty_and_all_out_of_range():
    v, c, n = run_c(np.zeros((0, 5), np.float32))
    assert v.shape == (0, 3, 5) and c.shape == (0, 3) and n.shape == (0,)
    far = np.full((7, 5), 100.0, np.float32)
    v, c, n = run_c(far)
    assert v.shape[0] == 0


def test_boundaries_and_order():
    f32 = np.float32
    pts = np.array([
        [-8.0, -8.0, -1.5, 1, 0],      # exactly lo  -> cell (0,0,0)
        [8.0, 0.0, 0.0, 2, 0],         # x == hi     -> c == grid -> dropped
        [7.99, 7.99, 1.49, 3, 0],      # last cell
        [0.0, 0.0, 1.5, 4, 0],         # z == hi     -> dropped (z IS tested by the voxelizer)
        [0.0, 0.0, -1.6, 5, 0],        # z < lo      -> dropped
        [-8.0, -8.0, -1.5, 6, 0],      # duplicate of point 0 -> same voxel, second slot
        [np.nan, 0.0, 0.0, 7, 0],      # NaN -> dropped
    ], dtype=f32)
    v, c, n = run_c(pts, k=2, mv=10)
    assert c.tolist() == [[0, 0, 0], [11, 31, 31]]          # (z,y,x), first-appearance order
    assert n.tolist() == [2, 1]
    assert v[0, 0, 3] == 1 and v[0, 1, 3] == 6 and v[1, 0, 3] == 3 and v[1, 1, 3] == 0


def test_max_points_keeps_first_k_in_input_order():
    pts = np.zeros((9, 4), np.float32)
    pts[:, 3] = np.arange(9)
    v, c, n = run_c(pts, k=4, mv=10)
    assert n.tolist() == [4] and v[0, :, 3].tolist() == [0, 1, 2, 3]


def test_max_voxels_continue_semantics():
    # three cells A,B,C; budget 2: C's points vanish, later points of A/B still join (spconv 2.x)
    A, B, C = [-7.9, -7.9, -1.4], [0.1, 0.1, 0.1], [5.1, 5.1, 1.1]
    seq = [A, B, C, A, C, B, B]
    pts = np.array([p + [i] for i, p in enumerate(seq)], np.float32)
    v, c, n = run_c(pts, k=5, mv=2)
    assert n.tolist() == [2, 3]
    assert v[0, :2, 3].tolist() == [0, 3] and v[1, :3, 3].tolist() == [1, 5, 6]
    v, c, n = run_c(pts, k=5, mv=2, legacy=True)   # 1.x: loop stops at the first refused voxel
    assert n.tolist() == [1, 1]


def test_mean_vfe_matches_torch_formula():
    import torch
    pts = rand_pts(7, 5000)
    v, c, n = run_c(pts, k=3, mv=4000)
    got = V.mean_vfe(v, n)
    tv, tn = torch.from_numpy(v), torch.from_numpy(n).float()
    ref = tv.sum(dim=1) / torch.clamp_min(tn.view(-1, 1), min=1.0)     # mean_vfe.py L26-28
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-6, atol=1e-7)
