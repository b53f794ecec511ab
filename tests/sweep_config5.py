"""sweep of point count x voxel size (one nuScenes-shaped frame resampled to N points).
Per point: K1 voxelizer (GPU, CUDA events; CPU = the oracle's C loop on one core, the reference's per-worker behaviour),
isolated SubMConv3d(64->64) and SparseConv3d(64->64, stride 2) forward on the resulting active set (bf16 tensor-core path).
Prints a markdown table (committed as profiles/r01_sweep.md).  Lives under tests/ because its CPU column runs the oracle
(test infrastructure); not collected by pytest, run as `python tests/sweep_config5.py` on a GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from toda_b200 import ops, synth
from toda_b200.spconv_compat import pytorch as sp
from oracle import voxelize as OV     # CPU arm of this measurement script only

dev = torch.device("cuda", 0)
sp.set_conv_precision("bf16")
cfg = synth.CONFIGS["nus_0075"]
pcr = cfg["pc_range"]
torch.manual_seed(0)
subm = sp.SubMConv3d(64, 64, 3, padding=1, bias=False, indice_key="s").to(dev)
down = sp.SparseConv3d(64, 64, 3, stride=2, padding=1, bias=False, indice_key="d").to(dev)


def gpu_time(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(1000000)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


print("| points | voxel xy (m) | voxels | voxelize GPU (ms) | voxelize CPU 1 core (ms) | SubM 64->64 fwd (ms) | pairs/row | SparseConv 64->64 s2 fwd (ms) | out rows |")
print("|---|---|---|---|---|---|---|---|---|")
for n_pts in (50000, 100000, 200000, 300000, 500000, 1000000):
    pts = synth.make_frame("nus_0075", 21, n_points=n_pts)
    for vxy in (0.05, 0.075, 0.1, 0.15, 0.2):
        vs = [vxy, vxy, 0.2]
        grid = synth.grid_size_xyz(pcr, vs)
        cap = 400000
        p = torch.from_numpy(pts).to(dev)
        offs = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=dev)
        out = {}

        def vox():
            out["v"] = ops.voxelize(p, offs, pcr, vs, 10, cap, order=ops.ORDER_CANONICAL, grid=grid)
        t_vox = gpu_time(vox)
        voxels, coords, num, _ = out["v"]
        gen = OV.Point2VoxelCPU3d(vsize_xyz=vs, coors_range_xyz=pcr, num_point_features=pts.shape[1],
                                  max_num_points_per_voxel=10, max_num_voxels=cap)
        gen.point_to_voxel(OV.from_numpy(pts))                 # (first call allocates the dense lookup table)
        t0 = time.perf_counter()
        gen.point_to_voxel(OV.from_numpy(pts))
        t_cpu = (time.perf_counter() - t0) * 1e3
        v = coords.shape[0]
        feats = torch.randn(v, 64, device=dev)
        shape = [int(grid[2]) + 1, int(grid[1]), int(grid[0])]
        x = sp.SparseConvTensor(feats, coords, shape, 1).canonical(assume_canonical=True)
        with torch.no_grad():
            y = subm(x)          # builds the rulebook once (indice_key)
            t_subm = gpu_time(lambda: subm(x))
            rb = x.indice_dict["s"][0]
            pairs = float((rb.nbr_fwd >= 0).sum().item()) / max(1, v)
            z = down(x)
            t_down = gpu_time(lambda: down(x))
        print("| %d | %.3f | %d | %.3f | %.1f | %.3f | %.1f | %.3f | %d |" % (pts.shape[0], vxy, v, t_vox, t_cpu, t_subm, pairs, t_down,
                                                                  z.features.shape[0]))
        del x, y, z, feats
