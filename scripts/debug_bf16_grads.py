"""Per-layer relative L2 difference between the bf16 tensor-core path and the fp32 FFMA path (one full frame)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toda_b200.pcdet_plugin as P
from toda_b200 import ops, synth
from toda_b200.spconv_compat import pytorch as G

class Cfg(dict): __getattr__ = dict.get
DEV = "cuda"
cfg = synth.CONFIGS["nus_0075"]
frames, collated = synth.make_batch("nus_0075", 1)
grid = synth.grid_size_xyz(cfg["pc_range"], cfg["voxel_size"])
pts = torch.from_numpy(collated).to(DEV)
offs = torch.tensor([0, collated.shape[0]], dtype=torch.int32, device=DEV)
v, c, n, _ = ops.voxelize(pts, offs, cfg["pc_range"], cfg["voxel_size"], 10, 120000, num_features=5, xyz_col=1, feat_col=1, order=ops.ORDER_CANONICAL)
vf = ops.mean_vfe(v, n)
hc = P.HeightCompression(Cfg(NUM_BEV_FEATURES=256))
res = {}
smooth = len(sys.argv) > 1 and sys.argv[1] == "smooth"
for mode in ("fp32", "bf16", "fp32b"):
    torch.manual_seed(666)
    net = P.VoxelResBackBone8x(Cfg(), 5, grid).to(DEV).train()
    G.set_conv_precision("fp32" if mode == "fp32b" else mode)
    x = vf.clone().requires_grad_(True)
    bd = hc(net({"voxel_features": x, "voxel_coords": c, "batch_size": 1, "voxel_coords_canonical": True}))
    sf = bd["spatial_features"]
    if smooth:
        loss = sf.square().mean()
    else:
        cot = torch.randn(sf.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
        loss = (sf * cot).sum()
    loss.backward()
    res[mode] = dict(sf=sf.detach(), dx=x.grad, grads={k: p.grad.detach().clone() for k, p in net.named_parameters()},
                     acts={k: t.features.detach() for k, t in bd["multi_scale_3d_features"].items()})
G.set_conv_precision("fp32")
rl2 = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
print("run-to-run fp32 vs fp32: sf %.2e dx %.2e" % (rl2(res["fp32b"]["sf"], res["fp32"]["sf"]), rl2(res["fp32b"]["dx"], res["fp32"]["dx"])))
print("sf %.3e dx %.3e" % (rl2(res["bf16"]["sf"], res["fp32"]["sf"]), rl2(res["bf16"]["dx"], res["fp32"]["dx"])))
for k in res["fp32"]["acts"]:
    print("act", k, "%.3e" % rl2(res["bf16"]["acts"][k], res["fp32"]["acts"][k]))
for k, g in res["fp32"]["grads"].items():
    print("%-28s %.3e  |g|=%.3e" % (k, rl2(res["bf16"]["grads"][k], g), float(g.norm())))
