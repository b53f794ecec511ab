"""Pipeline timeline of CTA 0 of the cp.async tcgen05 forward kernel on a real (synthetic-frame) active set with tile masks
(development aid).  usage: debug_timeline_real.py CIN COUT [level]   level 1 = stride-1 voxels, 2 = after one stride-2 conv"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TODA_TC_FEED"] = "cpasync"
import numpy as np, torch
from toda_b200 import ops, _C, synth
cin, cout = int(sys.argv[1]), int(sys.argv[2])
level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda", 0)
cfg = synth.CONFIGS["nus_0075"]
frames, collated = synth.make_batch("nus_0075", 2)
offs = torch.from_numpy(np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)).to(dev)
grid = synth.grid_size_xyz(cfg["pc_range"], cfg["voxel_size"])
_, coords, _, _ = ops.voxelize(torch.from_numpy(collated).to(dev), offs, cfg["pc_range"], cfg["voxel_size"], 10, 120000, xyz_col=1,
                               feat_col=1, num_features=5, order=ops.ORDER_CANONICAL, grid=grid)
shape = [int(grid[2]) + 1, int(grid[1]), int(grid[0])]
index = ops.OccupancyIndex(2, shape, dev, "dbg")
index.insert(coords); index.build(coords.shape[0], known_n=coords.shape[0])
for lv in range(1, level):
    rbs, index = ops.rulebook_sparse(index, [3, 3, 3], [2, 2, 2], [1, 1, 1], ("dbg", lv))
rb = ops.rulebook_subm(index, [3, 3, 3])
n = rb.n_in
x = torch.randn(n, cin, device=dev)
w = torch.randn(cout, 3, 3, 3, cin, device=dev) * 0.1
L = _C.lib()
buf = torch.zeros(8 * 256, dtype=torch.int64, device=dev)
for _ in range(3): ops.sparse_conv(x, w, None, rb, ops.CONV_BF16)
L.toda_debug_set_timeline.argtypes = [ctypes.c_void_p]
L.toda_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.sparse_conv(x, w, None, rb, ops.CONV_BF16); e1.record(); torch.cuda.synchronize()
L.toda_debug_set_timeline(None)
t = buf.cpu().numpy().reshape(8, 256).astype(np.int64)
masks = rb.tile_masks.cpu().numpy()
t0 = t[0, 0]
print("conv %d->%d level %d: n=%d, %.3f ms incl. pre-pass; pairs/row %.2f; tiles %d" % (cin, cout, level, n, e0.elapsed_time(e1),
      float((rb.nbr_fwd >= 0).sum().item()) / n, (n + 127) // 128))
print("tile  first_chunk  nchunks   start  slice_ready(+wait)  last_signalled  last_mma_done  drained   period")
prev = None
for it in range(4, 30):
    g0, g1 = int(t[7, it]), int(t[7, it + 1])
    if g1 <= g0 or g1 >= 256: break
    start, ready, drained = t[0, it] - t0, t[1, it] - t0, t[2, it] - t0
    last_sig = t[3, g1 - 1] - t0
    last_mma = t[6, g1 - 1] - t0
    print("%4d %8d %8d %9d %9d (+%5d) %12d %12d %10d %8s" % (it, g0, g1 - g0, start, ready, ready - start, last_sig, last_mma, drained,
          "" if prev is None else str(start - prev)))
    prev = start
