"""CPU analysis behind two tables of profiles/r01_feed_ab.md (no GPU needed):
  * non-empty K chunks per 128-row tile of the SubM tables (what the tile masks let the kernels skip), raster order vs
    rows sorted by chunk pattern;
  * gathered rows vs DISTINCT input rows per tile under raster, Morton and 4x4x8-block row orders (the case for a per-tile
    row cache).
One synthetic nuScenes-shaped bench frame, numpy only:   python scripts/analyze_tile_reuse.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from toda_b200 import synth  # noqa: E402


def level1_coords():
    cfg = synth.CONFIGS["nus_0075"]
    pts = np.asarray(synth.make_batch("nus_0075", 1)[1])
    r, vs = np.array(cfg["pc_range"], np.float32), np.array(cfg["voxel_size"], np.float32)
    c = np.floor((pts[:, 1:4] - r[:3]) / vs).astype(np.int64)
    grid = np.round((r[3:] - r[:3]) / vs).astype(np.int64)
    c = c[((c >= 0) & (c < grid)).all(1)]
    u = np.unique((c[:, 2] * grid[1] + c[:, 1]) * grid[0] + c[:, 0])
    coords = np.stack([u // (grid[1] * grid[0]), (u // grid[0]) % grid[1], u % grid[0]], 1)
    return coords, np.array([grid[2] + 1, grid[1], grid[0]])


def nbr_table(coords, shape):
    key = (coords[:, 0] * shape[1] + coords[:, 1]) * shape[2] + coords[:, 2]
    order = np.argsort(key)
    s = key[order]
    tbl = np.full((27, len(coords)), -1, np.int64)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                q = coords + np.array([dz, dy, dx])
                inside = ((q >= 0) & (q < shape)).all(1)
                kk = (q[:, 0] * shape[1] + q[:, 1]) * shape[2] + q[:, 2]
                idx = np.searchsorted(s, kk)
                idx[idx >= len(s)] = 0
                hit = inside & (s[idx] == kk)
                tbl[k, hit] = order[idx[hit]]
                k += 1
    return tbl


def downsample(coords, shape, k, s, p):
    out_shape = (shape + 2 * np.array(p) - np.array(k)) // np.array(s) + 1
    outs = []
    for dz in range(k[0]):
        for dy in range(k[1]):
            for dx in range(k[2]):
                o = coords + np.array(p) - np.array([dz, dy, dx])
                o = o[(o % np.array(s) == 0).all(1)] // np.array(s)
                o = o[((o >= 0) & (o < out_shape)).all(1)]
                outs.append((o[:, 0] * out_shape[1] + o[:, 1]) * out_shape[2] + o[:, 2])
    u = np.unique(np.concatenate(outs))
    return np.stack([u // (out_shape[1] * out_shape[2]), (u // out_shape[2]) % out_shape[1], u % out_shape[2]], 1), out_shape


def morton(coords):
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(12):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return out
    return (spread(coords[:, 0]) << np.uint64(2)) | (spread(coords[:, 1]) << np.uint64(1)) | spread(coords[:, 2])


def block_key(coords, bz, by, bx, shape):
    nb = ((shape[1] + by - 1) // by, (shape[2] + bx - 1) // bx)
    inner = ((coords[:, 0] % bz) * by + coords[:, 1] % by) * bx + coords[:, 2] % bx
    return (((coords[:, 0] // bz) * nb[0] + coords[:, 1] // by) * nb[1] + coords[:, 2] // bx) * (bz * by * bx) + inner


def analyse(name, coords, shape, offsets_per_chunk):
    tbl = nbr_table(coords, shape)
    n = len(coords)
    # --- chunks per tile
    nch = (27 + offsets_per_chunk - 1) // offsets_per_chunk
    cm = np.zeros(n, np.int64)
    for ch in range(nch):
        present = (tbl[ch * offsets_per_chunk:(ch + 1) * offsets_per_chunk] >= 0).any(0)
        cm |= present.astype(np.int64) << ch

    def chunks_per_tile(order):
        x = cm[order]
        x = np.concatenate([x, np.zeros((-len(x)) % 128, np.int64)]).reshape(-1, 128)
        return float(np.mean([bin(v).count("1") for v in np.bitwise_or.reduce(x, axis=1)]))
    print("%s: n=%d  K chunks per tile: %d total, %.2f non-empty in raster order, %.2f with rows sorted by chunk pattern"
          % (name, n, nch, chunks_per_tile(np.arange(n)), chunks_per_tile(np.argsort(cm, kind="stable"))))
    # --- reuse per tile
    for oname, order in (("raster", np.arange(n)), ("morton", np.argsort(morton(coords), kind="stable")),
                         ("4x4x8 blocks", np.argsort(block_key(coords, 4, 4, 8, shape), kind="stable"))):
        t = tbl[:, order]
        hits = distinct = tiles = 0
        for t0 in range(0, n, 128):
            v = t[:, t0:t0 + 128]
            v = v[v >= 0]
            hits += len(v)
            distinct += len(np.unique(v))
            tiles += 1
        print("    %-13s gathered rows / tile %.0f, distinct rows / tile %.0f (%.2f)" % (oname, hits / tiles, distinct / tiles,
                                                                                    distinct / hits))


if __name__ == "__main__":
    c, s = level1_coords()
    analyse("level 1 (16 ch)", c, s, 8)
    c, s = downsample(c, s, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    analyse("level 2 (32 ch)", c, s, 4)
    c, s = downsample(c, s, (3, 3, 3), (2, 2, 2), (1, 1, 1))
    analyse("level 3 (64 ch)", c, s, 2)
    c, s = downsample(c, s, (3, 3, 3), (2, 2, 2), (0, 1, 1))
    analyse("level 4 (128 ch)", c, s, 1)
