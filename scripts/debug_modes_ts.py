"""Ablation timing of the row-cache conv kernel (development aid): which role bounds it?  usage: debug_modes_ts.py CIN COUT level"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from toda_b200 import ops, _C, synth
cin, cout = int(sys.argv[1]), int(sys.argv[2])
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS["nus_0075"]
frames, collated = synth.make_batch("nus_0075", 4)
offs = torch.from_numpy(np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)).to(dev)
grid = synth.grid_size_xyz(cfg["pc_range"], cfg["voxel_size"])
_, coords, _, _ = ops.voxelize(torch.from_numpy(collated).to(dev), offs, cfg["pc_range"], cfg["voxel_size"], 10, 120000, xyz_col=1,
                               feat_col=1, num_features=5, order=ops.ORDER_CANONICAL, grid=grid)
shape = [int(grid[2]) + 1, int(grid[1]), int(grid[0])]
index = ops.OccupancyIndex(4, shape, dev, "dbg")
index.insert(coords); index.build(coords.shape[0], known_n=coords.shape[0])
for lv in range(1, level):
    rbs, index = ops.rulebook_sparse(index, [3, 3, 3], [2, 2, 2], [1, 1, 1], ("dbg", lv), cin=16, cout=16)
rb = ops.rulebook_subm(index, [3, 3, 3], channels=max(cin, cout))
n = rb.n_in
x = torch.randn(n, cin, device=dev)
xb = x.to(torch.bfloat16)
w = torch.randn(cout, 3, 3, 3, cin, device=dev) * 0.1
L = _C.lib()
L.toda_debug_set_mode.argtypes = [ctypes.c_int]
bits = ["no loads", "no STTM", "no MMA", "no slab copies", "no epi stores", "no slab protocol", "no lidx", "late slab release"]
print("conv %d->%d level %d n=%d tiles %d" % (cin, cout, level, n, (n + 127) // 128))
y = torch.empty((n, cout), device=dev)
wk = ops._repack(w, False, False)
modes = [int(m) for m in os.environ["TS_MODES"].split(",")] if os.environ.get("TS_MODES") else (0, 1, 2, 4, 8, 16, 32 | 8, 64, 31, 127, 0)
for mode in modes:
    L.toda_debug_set_mode(mode)
    for _ in range(2): ops._conv_call(x, xb, cin, rb.nbr_fwd, n, 27, wk, cout, None, ops.CONV_BF16, tile_masks=rb.tile_masks, plan=rb.plan, y=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5): ops._conv_call(x, xb, cin, rb.nbr_fwd, n, 27, wk, cout, None, ops.CONV_BF16, tile_masks=rb.tile_masks, plan=rb.plan, y=y)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print("mode %3d %-70s %.3f ms" % (mode, ", ".join(b for i, b in enumerate(bits) if mode >> i & 1) or "full", e0.elapsed_time(e1) / 5))
L.toda_debug_set_mode(0)
import os as _o
_o.environ["TODA_TC_FEED"] = "x"
