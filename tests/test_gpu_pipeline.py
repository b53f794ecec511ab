"""Input pipeline (side-stream front, planned geometry) against the in-line path: identical results, no index kernels in
the forward pass."""
import numpy as np
import pytest
import torch

from tests import parity_utils as PU

pytestmark = pytest.mark.gpu


def _modules(dev, precision):
    import toda_b200.pcdet_plugin as P
    from toda_b200 import synth
    from toda_b200.spconv_compat import pytorch as sp
    sp.set_conv_precision(precision)
    cfg = synth.CONFIGS["nus_0075"]
    pcr = [0.0, -9.6, -5.0, 19.2, 9.6, 3.0]                       # a 256 x 256 x 40 window of the nuScenes grid
    grid = synth.grid_size_xyz(pcr, cfg["voxel_size"])
    torch.manual_seed(666)
    vfe = P.MeanVFE(PU.Cfg(MAX_POINTS_PER_VOXEL=10, MAX_NUMBER_OF_VOXELS={"train": 20000, "test": 20000}), 5,
                    voxel_size=cfg["voxel_size"], point_cloud_range=pcr, grid_size=grid).to(dev)
    net = P.VoxelResBackBone8x(PU.Cfg(), 5, grid).to(dev).train()
    hc = P.HeightCompression(PU.Cfg(NUM_BEV_FEATURES=256))
    return vfe, net, hc, pcr


def _batches(pcr, n_batches, frames_per_batch=2):
    from toda_b200 import synth
    out = []
    for j in range(n_batches):
        frames = []
        for b in range(frames_per_batch):
            f = synth.make_frame("nus_0075", 10 * j + b)
            keep = (f[:, 0] > pcr[0] - 0.5) & (f[:, 0] < pcr[3] + 0.5) & (f[:, 1] > pcr[1] - 0.5) & (f[:, 1] < pcr[4] + 0.5)
            frames.append(f[keep])
        collated = np.concatenate([np.concatenate([np.full((f.shape[0], 1), b, np.float32), f], 1) for b, f in enumerate(frames)])
        offs = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)
        out.append((torch.from_numpy(collated), torch.from_numpy(offs)))
    return out


def _run(net, hc, bd):
    net.zero_grad(set_to_none=True)
    bd = hc(net(bd))
    sf = bd["spatial_features"]
    cot = torch.randn(sf.shape, device=sf.device, generator=torch.Generator(device=sf.device).manual_seed(3))
    (sf * cot).sum().backward()
    grads = {n: p.grad.detach().clone() for n, p in net.named_parameters()}
    return sf.detach().clone(), bd["encoded_spconv_tensor"].indices.clone(), grads


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pipelined_steps_equal_inline_steps(precision):
    from toda_b200 import ops
    from toda_b200.pipeline import InputPipeline
    from toda_b200.spconv_compat import pytorch as sp
    dev = torch.device("cuda", 0)
    try:
        vfe, net, hc, pcr = _modules(dev, precision)
        host = _batches(pcr, 3)
        # in-line reference results (BN running stats advance per step: replay from the same state afterwards)
        state = {k: v.clone() for k, v in net.state_dict().items()}
        want = []
        for pts, offs in host:
            bd = vfe({"points": pts.to(dev), "point_frame_offsets": offs.to(dev), "batch_size": 2})
            want.append(_run(net, hc, bd))
        net.load_state_dict(state)
        # pipelined: front of batch i+1 is submitted (from pinned host memory) while batch i runs
        pipe = InputPipeline(vfe, net, dev, reserve_bytes=64 << 20)
        pinned = [(p.pin_memory(), o.pin_memory()) for p, o in host]
        h = pipe.submit({"points": pinned[0][0], "point_frame_offsets": pinned[0][1], "batch_size": 2})
        for i in range(3):
            bd = pipe.consume(h)
            before = ops.launches()
            feats = bd["voxel_features"]
            assert bd["spconv_geometry"].coords.shape[0] == feats.shape[0]
            if i + 1 < 3:
                h = pipe.submit({"points": pinned[i + 1][0], "point_frame_offsets": pinned[i + 1][1], "batch_size": 2})
                before = ops.launches()
            got = _run(net, hc, bd)
            sf_w, idx_w, g_w = want[i]
            assert torch.equal(got[1], idx_w)
            if precision == "fp32":      # deterministic kernels: bit-identical
                assert torch.equal(got[0], sf_w), "spatial_features differ between the pipelined and the in-line step"
                for name in g_w:
                    assert torch.equal(got[2][name], g_w[name]), name
            else:                        # the fused BN statistics are accumulated with fp64 atomics (order-dependent last bits)
                PU.assert_close(got[0].cpu().numpy(), sf_w.cpu().numpy(), rtol=1e-4, what="spatial_features")
                for name in g_w:
                    PU.assert_close(got[2][name].cpu().numpy(), g_w[name].cpu().numpy(), rtol=1e-3, atol_scale=1e-4, what=name)
        torch.cuda.synchronize()
    finally:
        sp.set_conv_precision("fp32")


def test_planned_forward_has_no_geometry_launches():
    from toda_b200 import ops
    from toda_b200.spconv_compat import pytorch as sp
    dev = torch.device("cuda", 0)
    vfe, net, hc, pcr = _modules(dev, "fp32")
    pts, offs = _batches(pcr, 1)[0]
    bd = vfe({"points": pts.to(dev), "point_frame_offsets": offs.to(dev), "batch_size": 2})
    plan = net.plan_geometry(bd["voxel_coords"], 2, canonical=True)
    keys = set(plan.indice_dict)
    assert keys == {"subm1", "res1", "spconv2", "res2", "spconv3", "res3", "spconv4", "res4", "spconv_down2"}
    bd["spconv_geometry"] = plan
    ops.profile_begin()
    with torch.no_grad():
        net.eval()
        hc(net(bd))
    names = {r["name"] for r in ops.profile_end()}
    assert not (names & {"index_build", "rulebook_subm", "rulebook_sparse"}), names
    # a plan made for another batch is refused
    bd2 = dict(bd)
    bd2["voxel_features"] = bd["voxel_features"][:-1]
    with pytest.raises(ValueError):
        net(bd2)
