"""Full-size check of the row-cache conv kernel on every conv geometry of VoxelResBackBone8x (development aid):
plan kernel vs the round-1 kernel on the same tables, forward and dgrad."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from toda_b200 import ops, synth
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
chain = [(16, 32, [3, 3, 3], [2, 2, 2], [1, 1, 1]), (32, 64, [3, 3, 3], [2, 2, 2], [1, 1, 1]), (64, 128, [3, 3, 3], [2, 2, 2], [0, 1, 1]),
         (128, 128, [3, 1, 1], [2, 1, 1], [0, 0, 0])]
torch.manual_seed(0)
def check(name, x, cin, nbr, n_out, kvol, w, cout, plan, masks, out_rows=None):
    xb = x.to(torch.bfloat16)
    sums_a = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    sums_b = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    ya = ops._conv_call(x, xb, cin, nbr, n_out, kvol, w, cout, None, ops.CONV_BF16, tile_masks=masks, plan=None, bn_sums=sums_a, out_rows=out_rows)
    torch.cuda.synchronize()
    yb = ops._conv_call(x, xb, cin, nbr, n_out, kvol, w, cout, None, ops.CONV_BF16, tile_masks=masks, plan=plan, bn_sums=sums_b, out_rows=out_rows)
    torch.cuda.synchronize()
    err = float((ya - yb).abs().max()); ref = float(ya.abs().max())
    lid = plan.lidx.view(torch.int16).cpu().numpy().astype(np.uint16)
    print("%-22s n_in %7d n_out %7d  max|diff| %.3e (max|y| %.3e)  sums diff %.2e  fallback %.3f  blocks/unit mean %.1f max %d cap %d" %
          (name, x.shape[0], n_out, err, ref, float((sums_a - sums_b).abs().max() / sums_a.abs().max()), (lid == 0xFFFE).mean() / max((lid != 0xFFFF).mean(), 1e-9),
           float(plan.cnt.float().mean()), int(plan.cnt.max()), plan.cap))
for cin, cout, k, s, p in chain:
    rbs = ops.rulebook_subm(index, [3, 3, 3], channels=cin)
    x = torch.randn(rbs.n_in, cin, device=dev)
    w = torch.randn(27, cin, cin, device=dev) * 0.1
    check("subm %d->%d" % (cin, cin), x, cin, rbs.nbr_fwd, rbs.n_out, 27, w, cin, rbs.plan, rbs.tile_masks)
    rb, index = ops.rulebook_sparse(index, k, s, p, ("dbg", cin), cin=cin, cout=cout)
    kvol = k[0] * k[1] * k[2]
    w = torch.randn(kvol, cin, cout, device=dev) * 0.1
    check("down %d->%d fwd" % (cin, cout), x, cin, rb.nbr_fwd, rb.n_out, kvol, w, cout, rb.plan, rb.tile_masks)
    dy = torch.randn(rb.n_out, cout, device=dev)
    wt = torch.randn(kvol, cout, cin, device=dev) * 0.1
    if rb.dgrad_order is not None:
        check("down %d->%d dgrad" % (cin, cout), dy, cout, rb.nbr_bwd_sorted, rb.n_in, kvol, wt, cin, rb.dgrad_plan, rb.dgrad_tile_masks, out_rows=rb.dgrad_order)
    else:
        check("down %d->%d dgrad" % (cin, cout), dy, cout, rb.nbr_bwd, rb.n_in, kvol, wt, cin, rb.dgrad_plan, None)
print("ok")
