"""bf16 tensor-core (tcgen05) convolution path: parity against the CPU oracle evaluated on the SAME
bf16-rounded operands (then only the fp32 accumulation order differs: rtol 1e-4), plus the stated
end-to-end bf16 tolerance against the pure-fp32 oracle."""
import numpy as np
import pytest
import torch

from tests import parity_utils as PU

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF16_E2E_RTOL = 5e-2     # stated bf16 tolerance vs the fp32 oracle: relative L2 error of the encoder output


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("subm,cin,cout,k,s,p", [
    (True, 5, 16, 3, 1, 1), (True, 4, 16, 3, 1, 1), (True, 16, 16, 3, 1, 1), (True, 32, 32, 3, 1, 1), (True, 64, 64, 3, 1, 1), (True, 128, 128, 3, 1, 1),
    (False, 16, 32, 3, 2, 1), (False, 32, 64, 3, 2, 1), (False, 64, 128, 3, 2, (0, 1, 1)),
    (False, 128, 128, (3, 1, 1), (2, 1, 1), 0), (False, 64, 64, 3, 2, (0, 1, 1))])
def test_tc_conv_fwd_dgrad_vs_bf16_oracle(subm, cin, cout, k, s, p):
    from oracle import spconv_oracle as S
    from toda_b200.spconv_compat import pytorch as G
    torch.manual_seed(0)
    mk = (lambda sp: sp.SubMConv3d(cin, cout, k, padding=p, bias=True, indice_key="a")) if subm else \
        (lambda sp: sp.SparseConv3d(cin, cout, k, stride=s, padding=p, bias=True, indice_key="a"))
    a, b = mk(S), mk(G)
    with torch.no_grad():
        a.weight.copy_(bf16_round(a.weight))
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    shape, n, batch = [9, 24, 31], 3000, 2
    feats, idx = PU.random_sparse(7, batch, shape, n, cin)
    feats = bf16_round(torch.from_numpy(feats))
    fa = feats.clone().requires_grad_(True)
    fb = feats.clone().to(DEV).requires_grad_(True)
    ya = a(S.SparseConvTensor(fa, torch.from_numpy(idx), shape, batch))
    G.set_conv_precision("bf16")
    try:
        yb = b(G.SparseConvTensor(fb, torch.from_numpy(idx).to(DEV), shape, batch))
        assert np.array_equal(yb.indices.cpu().numpy(), ya.indices.numpy())
        PU.assert_close(yb.features.detach().cpu().numpy(), ya.features.detach().numpy(), what="tc fwd")
        g = bf16_round(torch.randn(ya.features.shape, generator=torch.Generator().manual_seed(3)))
        ya.features.backward(g)
        yb.features.backward(g.to(DEV))
    finally:
        G.set_conv_precision("fp32")
    PU.assert_close(fb.grad.cpu().numpy(), fa.grad.numpy(), what="tc dgrad")
    PU.assert_close(b.weight.grad.cpu().numpy(), a.weight.grad.numpy(), what="wgrad")
    PU.assert_close(b.bias.grad.cpu().numpy(), a.bias.grad.numpy(), what="bias grad")


def test_tc_ragged_tail_and_missing_neighbours():
    """n_out not a multiple of the 128-row tile; isolated voxels (all 26 neighbours missing)."""
    from oracle import spconv_oracle as S
    from toda_b200.spconv_compat import pytorch as G
    torch.manual_seed(1)
    a = S.SubMConv3d(32, 64, 3, padding=1, bias=False, indice_key="a")
    with torch.no_grad():
        a.weight.copy_(bf16_round(a.weight))
    b = G.SubMConv3d(32, 64, 3, padding=1, bias=False, indice_key="a")
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    for n in (1, 127, 129, 1000):
        feats, idx = PU.random_sparse(n, 1, [40, 50, 60], n, 32)
        feats = bf16_round(torch.from_numpy(feats))
        ya = a(S.SparseConvTensor(feats, torch.from_numpy(idx), [40, 50, 60], 1))
        G.set_conv_precision("bf16")
        try:
            yb = b(G.SparseConvTensor(feats.to(DEV), torch.from_numpy(idx).to(DEV), [40, 50, 60], 1))
        finally:
            G.set_conv_precision("fp32")
        PU.assert_close(yb.features.detach().cpu().numpy(), ya.features.detach().numpy(), what=f"tc n={n}")


def test_backbone_bf16_end_to_end_tolerance():
    """VoxelResBackBone8x on the golden crop in bf16 mode vs the fp32 reference golden: stated tolerance."""
    import toda_b200.pcdet_plugin as P
    from toda_b200.spconv_compat import pytorch as G
    g = PU.load_golden("backbone_res.npz")
    twin, net = PU.build_pair("VoxelResBackBone8x", 5, g["grid_size"], seed=int(g["seed"]))
    hc = P.HeightCompression(PU.Cfg(NUM_BEV_FEATURES=256))
    vf = torch.from_numpy(g["voxel_features"]).to(DEV)
    vc = torch.from_numpy(g["voxel_coords"]).float().to(DEV)
    cot = torch.randn(tuple(g["bev_shape"]), generator=torch.Generator().manual_seed(int(g["seed"])))
    G.set_conv_precision("bf16")
    try:
        r = PU.run_backbone(net, hc, vf, vc, 2, cot=cot, train=True)
    finally:
        G.set_conv_precision("fp32")
    f_sorted, i_sorted = PU.sort_rows(r["enc_features"], r["enc_indices"])
    assert np.array_equal(i_sorted, g["train_enc_indices"])           # geometry is precision-independent: bit-exact
    def rel_l2(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
    # stated bf16 tolerance (21 layers of bf16-rounded operands, fp32 accumulate, batch-statistics BN):
    # relative L2 error <= 5e-2 for activations, <= 1.5e-1 for the gradient that crossed the whole network backwards
    e_feat = rel_l2(f_sorted, g["train_enc_features"])
    e_grad = rel_l2(r["dvoxel_features"], g["train_dvoxel_features"])
    ga, gb = r["dvoxel_features"].ravel().astype(np.float64), g["train_dvoxel_features"].ravel().astype(np.float64)
    cos = float(ga @ gb / (np.linalg.norm(ga) * np.linalg.norm(gb)))
    print("bf16 end-to-end (golden crop): features rel-L2 %.3e, d voxel_features rel-L2 %.3e cos %.4f" % (e_feat, e_grad, cos))
    assert e_feat <= BF16_E2E_RTOL, e_feat
    # the crop is tiny (the last BatchNorms see ~250 rows), so batch-statistics BN backward amplifies the bf16 rounding
    # of 21 layers; the direction of the input gradient is what is asserted here, the full-size frame test below
    # bounds the magnitude error on a well-conditioned batch
    assert cos >= 0.85, cos
    names = [str(n) for n in g["grad_names"]]
    norms = np.array([np.linalg.norm(r["grads"][n].astype(np.float64)) for n in names])
    big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
    np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=0.25)


def test_bf16_vs_fp32_full_size_frame():
    """One full-size nuScenes-shaped frame (BASELINE configs[2] geometry): the tensor-core path against the fp32 FFMA
    path with identical weights.  Integer results are bit-identical; activations / gradients within the stated bf16
    tolerance (relative L2)."""
    import toda_b200.pcdet_plugin as P
    from toda_b200 import ops, synth
    from toda_b200.spconv_compat import pytorch as G
    cfg = synth.CONFIGS["nus_0075"]
    frames, collated = synth.make_batch("nus_0075", 1)
    grid = synth.grid_size_xyz(cfg["pc_range"], cfg["voxel_size"])
    pts = torch.from_numpy(collated).to(DEV)
    offs = torch.tensor([0, collated.shape[0]], dtype=torch.int32, device=DEV)
    v, c, n, _ = ops.voxelize(pts, offs, cfg["pc_range"], cfg["voxel_size"], cfg["max_points"], cfg["max_voxels"]["train"],
                              num_features=5, xyz_col=1, feat_col=1, order=ops.ORDER_CANONICAL)
    vf = ops.mean_vfe(v, n)
    hc = P.HeightCompression(PU.Cfg(NUM_BEV_FEATURES=256))
    res = {}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(666)
        net = P.VoxelResBackBone8x(PU.Cfg(), 5, grid).to(DEV).train()
        G.set_conv_precision(mode)
        try:
            x = vf.clone().requires_grad_(True)
            bd = hc(net({"voxel_features": x, "voxel_coords": c, "batch_size": 1, "voxel_coords_canonical": True}))
            sf = bd["spatial_features"]
            cot = torch.randn(sf.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
            (sf * cot).sum().backward()
        finally:
            G.set_conv_precision("fp32")
        res[mode] = dict(sf=sf.detach(), idx=bd["encoded_spconv_tensor"].indices, dx=x.grad,
                         grads={k: p.grad.detach() for k, p in net.named_parameters()})

    def rl2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    assert torch.equal(res["fp32"]["idx"], res["bf16"]["idx"])

    def cos(a, b):
        a, b = a.double().flatten(), b.double().flatten()
        return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))
    e_sf = rl2(res["bf16"]["sf"], res["fp32"]["sf"])
    c_dx = cos(res["bf16"]["dx"], res["fp32"]["dx"])
    c_w = {k: cos(res["bf16"]["grads"][k], g) for k, g in res["fp32"]["grads"].items() if k.endswith("weight") and g.dim() == 5}
    print("bf16 vs fp32, full frame: spatial_features rel-L2 %.3e; cosine d voxel_features %.4f, wgrad min %.4f median %.4f"
          % (e_sf, c_dx, min(c_w.values()), float(np.median(list(c_w.values())))))
    # Forward: stated bf16 tolerance.  Backward: with a noise cotangent every gradient is a random-walk sum over ~1e5
    # rows, so the ~1 % of ReLU masks that flip under a 1 % forward perturbation move it by sqrt(1 %) ~ 10-30 % in L2
    # whatever the arithmetic (scripts/debug_bf16_grads.py shows 9 % already on the last layer's BN bias, which depends
    # on the masks only).  Kernel-level backward parity is asserted exactly elsewhere (same inputs: rtol 1e-3/1e-4);
    # end to end the direction of the gradients is what is bounded.
    assert e_sf <= BF16_E2E_RTOL
    assert c_dx >= 0.9 and min(c_w.values()) >= 0.9


def test_bf16_gradients_with_frozen_masks_full_size_frame():
    """The numeric gradient bound of the bf16 path: one full-size nuScenes-shaped frame, fp32 FFMA path vs bf16 tcgen05 path
    with identical weights AND identical ReLU masks (recorded in the fp32 pass, imposed on the bf16 pass with
    ops.ReluMaskTape), so that what is compared is the arithmetic (bf16-rounded operands, fp32 accumulation, 21 layers,
    batch-statistics BatchNorm), not the mask flips a perturbed forward pass causes.  Stated tolerance: relative L2 of
    d loss / d voxel_features and of every conv weight gradient <= 3e-2 (measured: see the printed line)."""
    import toda_b200.pcdet_plugin as P
    from toda_b200 import ops, synth
    from toda_b200.spconv_compat import pytorch as G
    cfg = synth.CONFIGS["nus_0075"]
    frames, collated = synth.make_batch("nus_0075", 1)
    grid = synth.grid_size_xyz(cfg["pc_range"], cfg["voxel_size"])
    pts = torch.from_numpy(collated).to(DEV)
    offs = torch.tensor([0, collated.shape[0]], dtype=torch.int32, device=DEV)
    v, c, n, _ = ops.voxelize(pts, offs, cfg["pc_range"], cfg["voxel_size"], cfg["max_points"], cfg["max_voxels"]["train"],
                              num_features=5, xyz_col=1, feat_col=1, order=ops.ORDER_CANONICAL)
    vf = ops.mean_vfe(v, n)
    hc = P.HeightCompression(PU.Cfg(NUM_BEV_FEATURES=256))
    tape = ops.ReluMaskTape("record")
    res = {}
    try:
        for mode in ("fp32", "bf16"):
            torch.manual_seed(666)
            net = P.VoxelResBackBone8x(PU.Cfg(), 5, grid).to(DEV).train()
            G.set_conv_precision(mode)
            ops.set_relu_mask_tape(tape)
            x = vf.clone().requires_grad_(True)
            bd = hc(net({"voxel_features": x, "voxel_coords": c, "batch_size": 1, "voxel_coords_canonical": True}))
            sf = bd["spatial_features"]
            cot = torch.randn(sf.shape, device=DEV, generator=torch.Generator(device=DEV).manual_seed(5))
            (sf * cot).sum().backward()
            res[mode] = dict(sf=sf.detach(), dx=x.grad, grads={k: p.grad.detach() for k, p in net.named_parameters()})
            tape.replay()
    finally:
        ops.set_relu_mask_tape(None)
        G.set_conv_precision("fp32")

    def rl2(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    e_sf = rl2(res["bf16"]["sf"], res["fp32"]["sf"])
    e_dx = rl2(res["bf16"]["dx"], res["fp32"]["dx"])
    e_w = {k: rl2(res["bf16"]["grads"][k], g) for k, g in res["fp32"]["grads"].items() if g.dim() == 5}
    e_bn = {k: rl2(res["bf16"]["grads"][k], g) for k, g in res["fp32"]["grads"].items() if g.dim() == 1 and ".bn" in k or k.endswith(".1.weight") or k.endswith(".1.bias")}
    print("bf16 vs fp32, masks frozen, full frame: spatial_features rel-L2 %.3e; d voxel_features %.3e; conv weight grads max %.3e "
          "median %.3e; BN affine grads max %.3e" % (e_sf, e_dx, max(e_w.values()), float(np.median(list(e_w.values()))),
                                                      max(e_bn.values())))
    assert e_sf <= 3e-2 and e_dx <= 3e-2
    assert max(e_w.values()) <= 3e-2, max(e_w.items(), key=lambda kv: kv[1])
    assert max(e_bn.values()) <= 3e-2, max(e_bn.items(), key=lambda kv: kv[1])


@pytest.mark.parametrize("cin,cout,n", [(16, 16, 300000), (32, 32, 200000), (64, 64, 150000), (128, 128, 60000), (64, 128, 90000)])
def test_tc_matches_ffma_path_at_scale(cin, cout, n):
    """Many tiles per persistent CTA, ring / phase wrap-around, TMEM double buffering, split-K wgrad over thousands of
    rows: the tensor-core kernels against the FFMA kernels on bf16-representable operands (same tables)."""
    from toda_b200 import ops
    rng = np.random.default_rng(0)
    shape, batch = [21, 400, 400], 2
    feats, idx = PU.random_sparse(3, batch, shape, n, cin)
    x = bf16_round(torch.from_numpy(feats)).to(DEV)
    index = ops.OccupancyIndex(batch, shape, torch.device(DEV, 0), "scale")
    index.insert(torch.from_numpy(idx).to(DEV))
    index.build(n)
    rb = ops.rulebook_subm(index, [3, 3, 3])
    w = bf16_round(torch.from_numpy(rng.standard_normal((cout, 3, 3, 3, cin)).astype(np.float32) * 0.1)).to(DEV)
    g = bf16_round(torch.from_numpy(rng.standard_normal((n, cout)).astype(np.float32))).to(DEV)
    out = {}
    for prec in (ops.CONV_FP32, ops.CONV_BF16):
        xx = x.clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        y = ops.sparse_conv(xx, ww, None, rb, prec)
        y.backward(g)
        out[prec] = (y.detach(), xx.grad, ww.grad)
    index.release()
    for name, a, b in zip(("fwd", "dgrad", "wgrad"), out[ops.CONV_BF16], out[ops.CONV_FP32]):
        PU.assert_close(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-3, atol_scale=1e-4, what=f"{name} {cin}->{cout} n={n}")


# ------------------------------------------------------------------------------------------ bf16x3: fp32-class accuracy on tensor cores
@pytest.mark.parametrize("subm,cin,cout,k,s,p", [
    (True, 5, 16, 3, 1, 1), (True, 4, 16, 3, 1, 1), (True, 16, 16, 3, 1, 1), (True, 32, 32, 3, 1, 1), (True, 64, 64, 3, 1, 1), (True, 128, 128, 3, 1, 1),
    (False, 16, 32, 3, 2, 1), (False, 64, 128, 3, 2, (0, 1, 1)), (False, 128, 128, (3, 1, 1), (2, 1, 1), 0)])
def test_x3_conv_matches_fp32_oracle_at_rtol_1e4(subm, cin, cout, k, s, p):
    """TODA_CONV_BF16X3 (split-bf16, three tcgen05 launches per product) against the fp32 oracle on UNROUNDED operands:
    forward, dgrad, wgrad and bias gradient within the north star's fp32 tolerance (rtol 1e-4)."""
    from oracle import spconv_oracle as S
    from toda_b200.spconv_compat import pytorch as G
    torch.manual_seed(0)
    mk = (lambda sp: sp.SubMConv3d(cin, cout, k, padding=p, bias=True, indice_key="a")) if subm else \
        (lambda sp: sp.SparseConv3d(cin, cout, k, stride=s, padding=p, bias=True, indice_key="a"))
    a, b = mk(S), mk(G)
    b.load_state_dict(a.state_dict())
    b = b.to(DEV)
    shape, n, batch = [9, 24, 31], 3000, 2
    feats, idx = PU.random_sparse(7, batch, shape, n, cin)
    fa = torch.from_numpy(feats).requires_grad_(True)
    fb = torch.from_numpy(feats).to(DEV).requires_grad_(True)
    ya = a(S.SparseConvTensor(fa, torch.from_numpy(idx), shape, batch))
    G.set_conv_precision("bf16x3")
    try:
        yb = b(G.SparseConvTensor(fb, torch.from_numpy(idx).to(DEV), shape, batch))
        assert np.array_equal(yb.indices.cpu().numpy(), ya.indices.numpy())
        PU.assert_close(yb.features.detach().cpu().numpy(), ya.features.detach().numpy(), what="x3 fwd")
        g = torch.randn(ya.features.shape, generator=torch.Generator().manual_seed(3))
        ya.features.backward(g)
        yb.features.backward(g.to(DEV))
    finally:
        G.set_conv_precision("fp32")
    PU.assert_close(fb.grad.cpu().numpy(), fa.grad.numpy(), what="x3 dgrad")
    PU.assert_close(b.weight.grad.cpu().numpy(), a.weight.grad.numpy(), what="x3 wgrad")
    PU.assert_close(b.bias.grad.cpu().numpy(), a.bias.grad.numpy(), what="x3 bias grad")


def test_backbone_x3_vs_reference_golden():
    """VoxelResBackBone8x in bf16x3 mode.  Forward: against the golden produced through the reference's own
    spconv_backbone.py.  Every single convolution meets rtol 1e-4 (tests above); through 21 layers with batch-statistics
    BatchNorm the ~1e-5 per-layer error compounds to ~1e-4, so the end-to-end tolerance stated for this mode is rtol 1e-3 on
    activations (measured: max error 4.1e-4 at scale 4.5; the bf16 mode's is 5e-2, the FFMA path's 1e-4).
    Backward: a 1e-5 perturbation is enough to flip ReLU masks of activations that sit at the kink, and one flipped mask
    moves the whole gradient field by ~0.5 % (measured on the stage-2 golden, tests/test_gpu_parity.py), so the gradient
    ARITHMETIC is compared with the masks frozen (recorded in an fp32 FFMA pass, imposed on the bf16x3 pass): relative L2
    <= 5e-4 for d loss / d voxel_features and every conv-weight / BatchNorm gradient (measured 2.6e-5 and 9.5e-5; the bf16
    mode's bound is 3e-2); against the
    golden, with its own masks, the gradients are held to relative L2 <= 2e-2."""
    import toda_b200.pcdet_plugin as P
    from toda_b200 import ops
    from toda_b200.spconv_compat import pytorch as G
    g = PU.load_golden("backbone_res.npz")
    hc = P.HeightCompression(PU.Cfg(NUM_BEV_FEATURES=256))
    vf = torch.from_numpy(g["voxel_features"]).to(DEV)
    vc = torch.from_numpy(g["voxel_coords"]).float().to(DEV)
    cot = torch.randn(tuple(g["bev_shape"]), generator=torch.Generator().manual_seed(int(g["seed"])))

    def rl2(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

    # (1) the product path (fused layers, own masks) against the golden
    twin, net = PU.build_pair("VoxelResBackBone8x", 5, g["grid_size"], seed=int(g["seed"]))
    G.set_conv_precision("bf16x3")
    try:
        r = PU.run_backbone(net, hc, vf, vc, 2, cot=cot, train=True)
    finally:
        G.set_conv_precision("fp32")
    f_sorted, i_sorted = PU.sort_rows(r["enc_features"], r["enc_indices"])
    assert np.array_equal(i_sorted, g["train_enc_indices"])
    PU.assert_close(f_sorted, g["train_enc_features"], rtol=1e-3, atol_scale=1e-4, what="x3 encoded features")
    assert rl2(r["dvoxel_features"], g["train_dvoxel_features"]) <= 2e-2
    assert rl2(r["grads"]["conv_out.0.weight"], g["grad_conv_out_weight"]) <= 2e-2
    assert rl2(r["grads"]["conv_input.0.weight"], g["grad_conv_input_weight"]) <= 2e-2

    # (2) gradient arithmetic with frozen ReLU masks: fp32 FFMA pass records, bf16x3 pass replays
    tape = ops.ReluMaskTape("record")
    res = {}
    try:
        for mode in ("fp32", "bf16x3"):
            twin, net = PU.build_pair("VoxelResBackBone8x", 5, g["grid_size"], seed=int(g["seed"]))
            G.set_conv_precision(mode)
            ops.set_relu_mask_tape(tape)
            res[mode] = PU.run_backbone(net, hc, vf, vc, 2, cot=cot, train=True)
            tape.replay()
    finally:
        ops.set_relu_mask_tape(None)
        G.set_conv_precision("fp32")
    e_dx = rl2(res["bf16x3"]["dvoxel_features"], res["fp32"]["dvoxel_features"])
    # (the bias of a convolution that feeds a batch-statistics BatchNorm has an identically zero gradient: both passes hold
    # rounding noise there, which has no relative error to speak of)
    e_p = {k: rl2(res["bf16x3"]["grads"][k], v) for k, v in res["fp32"]["grads"].items()
           if v.ndim == 5 or ".bn" in k or k.endswith(".1.weight") or k.endswith(".1.bias")}
    print("bf16x3 vs fp32, masks frozen: d voxel_features rel-L2 %.3e; parameter grads max %.3e" % (e_dx, max(e_p.values())))
    assert e_dx <= 5e-4          # measured 2.6e-5
    assert max(e_p.values()) <= 5e-4, max(e_p.items(), key=lambda kv: kv[1])    # measured 9.5e-5
