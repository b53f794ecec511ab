"""K0 (toda_points_select) timing on the bench batch: CUDA events, L2 flushed between iterations; prints a markdown table
(committed as profiles/r01_points_select.md)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from toda_b200 import ops, synth

dev = torch.device("cuda", 0)
frames, collated = synth.make_batch("nus_0075", 4)
offs = torch.from_numpy(np.cumsum([0] + [f.shape[0] for f in frames]).astype(np.int32)).to(dev)
pts = torch.from_numpy(collated).to(dev)
raw = torch.from_numpy(np.concatenate(frames)).to(dev)
peak = 6545.6
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = synth.CONFIGS["nus_0075"]["pc_range"]
cases = [
    ("range mask (keeps all: frames are pre-masked)", pts, offs, ops.SELECT_RANGE_XY, [r[0], r[1], r[3], r[4]], False, False, 1),
    ("range mask, half range", pts, offs, ops.SELECT_RANGE_XY, [r[0] / 2, r[1] / 2, r[3] / 2, r[4] / 2], False, False, 1),
    ("collate (raw frames -> [b,x,y,z,..])", raw, offs, ops.SELECT_RANGE_XY, [-1e30, -1e30, 1e30, 1e30], False, True, 0),
    ("cutmix rectangle, inside", pts, None, ops.SELECT_RECT_XY, [-20.0, -30.0, 35.0, 25.0], False, False, 1),
    ("PolarMix sector, complement", pts, None, ops.SELECT_SECTOR, [0.3, 0.3 + 1.570796], True, False, 1),
]
print("| case | points in | points out | us | algorithmic MB | GB/s | share of HBM peak (%.0f GB/s) |" % peak)
print("|---|---|---|---|---|---|---|")
for name, p, o, mode, params, inv, addb, xc in cases:
    ts = []
    for it in range(12):
        flush.zero_()
        torch.cuda._sleep(2000000)     # ~1 ms of GPU spin so that the host has queued the call before e0 is reached
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, oo = ops.points_select(p, o, mode, params, invert=inv, add_batch_col=addb, x_col=xc, trim=False)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    kept = int(oo[-1].item())
    us = float(np.median(ts))
    mb = (4.0 * p.shape[1] * p.shape[0] + 4.0 * out.shape[1] * kept) / 1e6
    gbs = mb / us * 1e3
    print("| %s | %d | %d | %.1f | %.1f | %.0f | %.1f%% |" % (name, p.shape[0], kept, us, mb, gbs, 100 * gbs / peak))
