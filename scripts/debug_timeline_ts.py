"""Pipeline timeline of CTA 0 of the row-cache / TMEM-operand conv kernel (conv_ts.cu) on a real (synthetic-frame) active
set (development aid).  usage: debug_timeline_ts.py CIN COUT [level]"""
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
w = torch.randn(cout, 3, 3, 3, cin, device=dev) * 0.1
L = _C.lib()
R = 24
buf = torch.zeros(R * 256, dtype=torch.int64, device=dev)
L.toda_debug_set_mode.argtypes = [ctypes.c_int]
L.toda_debug_set_mode(int(os.environ.get('TS_MODE', '0')))
for _ in range(3): ops.sparse_conv(x, w, None, rb, ops.CONV_BF16)
L.toda_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.toda_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.sparse_conv(x, w, None, rb, ops.CONV_BF16); e1.record(); torch.cuda.synchronize()
L.toda_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(R, 256).astype(np.int64)
cnt = rb.plan.cnt.cpu().numpy()
print("conv %d->%d level %d: n=%d, %.3f ms incl. pre-pass; pairs/row %.2f; tiles %d; mean distinct rows per (tile,group) %.1f max %d cap %d"
      % (cin, cout, level, n, e0.elapsed_time(e1), float((rb.nbr_fwd >= 0).sum().item()) / n, (n + 127) // 128, cnt.mean(), cnt.max(), rb.plan.cap))
lid = rb.plan.lidx.cpu().numpy().astype(np.uint16)
print("table entries: none %.3f  global-fallback %.5f" % ((lid == 0xFFFF).mean(), (lid == 0xFFFE).mean()))
t0 = t[7, 0]
tiles_g = t[10]
print("tile starts (g):", tiles_g[:16])
print("-- MMA warp per tile: start, +wait acc_empty")
for it in range(3, 12):
    print("tile %2d  start %8d  acc_empty wait %5d" % (it, int(t[14, it] - t0), int(t[15, it] - t[14, it])))
print("-- gather warp 8 per tile: loop top, +lidx wait | pairs done, +release all")
for it in range(3, 12):
    print("tile %2d  top %8d +%5d | pairs done %8d +%5d" % (it, int(t[22, it] - t0), int(t[23, it] - t[22, it]), int(t[20, it] - t0), int(t[21, it] - t[20, it])))
print("-- per pair (first g of the pair): gather: start, +release/slabwait, +loads issued & a_empty ok, +fix/sync, +sttm..arrive | MMA: start, +waits, +issue")
g_lo, g_hi = int(tiles_g[4]), int(tiles_g[9])
for g in range(g_lo, min(g_hi, 255)):
    if t[0, g] == 0:
        continue
    a = [int(t[r, g] - t0) for r in (0, 1, 12, 2, 3)]
    m = [int(t[r, g] - t0) for r in (4, 5)]
    wg = 0 if t[16, g] else 4
    arr = [int(t[16 + wg + i, g] - t0) for i in range(4)]
    print("g %3d  gather %8d +%5d +%5d +%5d +%5d = %5d | arrivals of the 4 warps %s | mma issue %8d +%5d" %
          (g, a[0], a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3], a[4] - a[0], arr, m[0], m[1] - m[0]))
print("-- loader per unit u: start | ids+slab_empty | issued | rows")
for u in range(12, 30):
    print("u %3d  %8d +%5d +%5d  R=%d" % (u, int(t[7, u] - t0), int(t[8, u] - t[7, u]), int(t[9, u] - t[8, u]), int(t[11, u])))
