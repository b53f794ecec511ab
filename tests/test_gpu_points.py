"""K0 point selection and the GPU point preprocessor / mixers against the reference-generated goldens
(tests/golden/points.npz) and the oracle (oracle/points.py) at full frame size.  Results must be bit-identical."""
import os

import numpy as np
import pytest
import torch

from oracle import points as OP
from tests import parity_utils as PU

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "points.npz"))
DEV = "cuda:0"


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    return (t.to(dtype) if dtype else t).to(DEV)


def test_range_mask_and_shuffle_golden():
    from toda_b200.pcdet_plugin.processor import PointPreprocessor
    pts, pcr = G["mask/points"], G["mask/pc_range"]
    cfgs = [PU.Cfg(NAME="mask_points_and_boxes_outside_range", REMOVE_OUTSIDE_BOXES=True),
            PU.Cfg(NAME="shuffle_points", SHUFFLE_ENABLED=PU.Cfg(train=True, test=True)),
            PU.Cfg(NAME="transform_points_to_voxels_placeholder", VOXEL_SIZE=[0.1, 0.1, 0.2])]
    pp = PointPreprocessor(cfgs, pcr, training=True, num_point_features=pts.shape[1])
    assert pp.grid_size.tolist() == [400, 240, 40]
    collated = np.pad(pts, ((0, 0), (1, 0)))
    dd = pp.forward({"points": _t(collated), "point_frame_offsets": _t(np.array([0, pts.shape[0]], np.int32)),
                     "shuffle_idx": G["mask/shuffle_idx"].astype(np.int32)})
    got = dd["points"].cpu().numpy()
    assert np.array_equal(got[:, 1:], G["mask/result"], equal_nan=True)
    assert dd["point_frame_offsets"].cpu().tolist() == [0, G["mask/result"].shape[0]]


def test_cutmix_golden():
    from toda_b200.pcdet_plugin.processor import cutmix_points
    got = cutmix_points(_t(G["cutmix/source"]), _t(G["cutmix/target"]), G["cutmix/min_xy"], G["cutmix/max_xy"])
    assert np.array_equal(got.cpu().numpy(), G["cutmix/result"])


def test_polar_swap_golden():
    from toda_b200.pcdet_plugin.processor import polar_swap_points
    for j in range(2):
        a0, a1 = G[f"polar{j}/angles"]
        got = polar_swap_points(_t(G[f"polar{j}/pt1"]), _t(G[f"polar{j}/pt2"]), a0, a1)
        assert np.array_equal(got.cpu().numpy(), G[f"polar{j}/result"])


def test_mixup_golden():
    from toda_b200.pcdet_plugin.processor import mixup_points
    got = mixup_points(_t(G["mixup/points_1"]), _t(G["mixup/points_2"]), float(G["mixup/lam"]), G["mixup/shuffle_idx_1"],
                       G["mixup/shuffle_idx_2"])
    assert np.array_equal(got.cpu().numpy(), G["mixup/result"])


def test_collate_golden():
    from toda_b200.pcdet_plugin.processor import collate_points
    frames = [G[f"collate/frame{i}"] for i in range(3)]
    raw = np.concatenate(frames)
    offs = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)
    got = collate_points(_t(raw), _t(offs))
    assert np.array_equal(got.cpu().numpy(), G["collate/result"])


@pytest.mark.parametrize("config", ["nus_0075", "waymo_010"])
def test_range_mask_full_batch_vs_oracle(config):
    """4 full-size frames, per-frame offsets, tile-boundary and empty-frame handling."""
    from toda_b200 import ops, synth
    cfg = synth.CONFIGS[config]
    frames = [synth.make_frame(config, i, shuffle=True) for i in range(3)]
    frames.insert(2, frames[0][:0])                               # an empty frame in the middle
    tight = np.array(cfg["pc_range"], np.float32) * np.float32(0.5)
    collated = OP.collate_points(frames).astype(np.float32)
    offs = np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)
    got, got_offs = ops.points_select(_t(collated), _t(offs), ops.SELECT_RANGE_XY, [tight[0], tight[1], tight[3], tight[4]], x_col=1)
    want = [OP.mask_points_outside_range(f, tight) for f in frames]
    assert np.array_equal(got.cpu().numpy(), OP.collate_points(want).astype(np.float32))
    assert got_offs.cpu().tolist() == np.cumsum([0] + [w.shape[0] for w in want]).tolist()


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1023, 1024, 1025, 4096, 70001])
def test_select_sizes_and_inversion(n):
    from toda_b200 import ops
    rng = np.random.default_rng(n)
    pts = (rng.standard_normal((n, 4)) * 10).astype(np.float32)
    for mode, params, oracle in [
        (ops.SELECT_RANGE_XY, [-5.0, -3.0, 7.0, 9.0], lambda p: OP.mask_points_by_range(p, [-5.0, -3.0, 0, 7.0, 9.0, 0])),
        (ops.SELECT_RECT_XY, [-5.25, -3.1, 7.3, 9.9], lambda p: OP.rect_mask(p, [-5.25, -3.1], [7.3, 9.9])),
        (ops.SELECT_SECTOR, [-1.0, 0.75], lambda p: OP.sector_mask(p, -1.0, 0.75)),
        (ops.SELECT_SECTOR, [0.5, 2.5, 8.0, 1.0], lambda p: OP.sector_mask(p, 0.5, 2.5, np.float32(8.0), True)),
        (ops.SELECT_SECTOR, [0.5, 2.5, 8.0, 2.0], lambda p: OP.sector_mask(p, 0.5, 2.5, np.float32(8.0), False)),
    ]:
        m = oracle(pts)
        for invert in (False, True):
            got, total = ops.points_select(_t(pts), None, mode, params, invert=invert)
            want = pts[~m] if invert else pts[m]
            assert int(total.item()) == want.shape[0], (n, mode, invert)
            assert np.array_equal(got.cpu().numpy(), want), (n, mode, invert)


def test_sector_full_frame_vs_oracle():
    """268 k points: membership of the -arctan2 sector must match numpy's float32 arctan2 point for point."""
    from toda_b200 import ops, synth
    pts = synth.make_frame("nus_0075", 11)
    for a0, a1 in [(-np.pi, -np.pi + 1.570796), (0.3, 0.3 + 1.570796), (2.2, np.pi)]:
        got, _ = ops.points_select(_t(pts), None, ops.SELECT_SECTOR, [a0, a1])
        assert np.array_equal(got.cpu().numpy(), pts[OP.sector_mask(pts, a0, a1)])


def test_points_in_box_golden():
    """The box test of the mixers (augmentor_utils.get_points_in_box, run by the reference itself for the golden): per box
    the selected rows are exactly points[mask]; all boxes together = the point filter of intra_domain_point_mixup_cd."""
    from toda_b200.pcdet_plugin.processor import points_in_boxes, remove_points_in_boxes
    pts, boxes, masks = G["inbox/points"], G["inbox/boxes"], G["inbox/masks"]
    tp = _t(pts)
    for b in range(boxes.shape[0]):
        got = points_in_boxes(tp, boxes[b:b + 1]).cpu().numpy()
        assert np.array_equal(got, pts[masks[b]]), b
    got = points_in_boxes(tp, boxes).cpu().numpy()
    assert np.array_equal(got, pts[masks.any(axis=0)])
    got = remove_points_in_boxes(tp, boxes).cpu().numpy()
    assert np.array_equal(got, pts[~masks.any(axis=0)])
    assert np.array_equal(got, OP.remove_points_in_boxes(pts, boxes))


def test_points_in_boxes_full_frame_vs_oracle():
    """A full-size frame against 150 random boxes (two chunks of the 96-box kernel limit), bit-identical to the oracle."""
    from toda_b200 import synth
    from toda_b200.pcdet_plugin.processor import remove_points_in_boxes
    pts = synth.make_frame("nus_0075", 3)
    rng = np.random.default_rng(7)
    n_box = 150
    centers = pts[rng.integers(0, pts.shape[0], n_box), :3] + rng.normal(0, 0.3, (n_box, 3)).astype(np.float32)
    sizes = rng.uniform(0.5, 6.0, (n_box, 3)).astype(np.float32)
    rz = rng.uniform(-np.pi, np.pi, (n_box, 1)).astype(np.float32)
    boxes = np.concatenate([centers.astype(np.float32), sizes, rz], axis=1).astype(np.float32)
    got = remove_points_in_boxes(_t(pts), boxes).cpu().numpy()
    ref = OP.remove_points_in_boxes(pts, boxes)
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert 0 < ref.shape[0] < pts.shape[0]
