"""Seeded synthetic LiDAR frames shaped like the datasets the reference trains on.

There is no dataset access, so throughput and parity are measured on ray-cast frames whose
point counts, feature columns and spatial statistics follow the reference's loaders:
  * nuScenes-shaped: 32 beams, 10 sweeps, F=5 [x,y,z,intensity,dt], ego-box removal |x|<1 & |y|<1
    (pcdet/datasets/nuscenes/nuscenes_dataset.py L82-115), sensor-centred frame.
  * Waymo-shaped: 64 beams x 2650 azimuth steps + short-range side lidars, F=5
    [x,y,z,tanh(intensity),elongation] (pcdet/datasets/waymo/waymo_dataset.py L159-167), vehicle frame.
Seeds: numpy.random.default_rng(1000*config_id + frame_idx).  The
`shuffle_points` permutation (data_processor.py L93-103) is drawn from the same generator and applied
here, so CPU and GPU paths see identical point order.
"""
import numpy as np

# workload constants restated from the reference's YAML configs
CONFIGS = {
    # cfg #1: tools/cfgs/dataset_configs/nuscenes_dataset.yaml L5, L20, L57-61, L74-80
    "nus_010": dict(kind="nus", pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], voxel_size=[0.1, 0.1, 0.2],
                    max_points=10, max_voxels={"train": 60000, "test": 60000}, num_features=5, config_id=1),
    # cfg #2: tools/cfgs/dataset_configs/waymo_dataset.yaml L5, L57, L73-78
    "waymo_010": dict(kind="waymo", pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15],
                      max_points=5, max_voxels={"train": 150000, "test": 150000}, num_features=5, config_id=2),
    # cfg #3: tools/cfgs/nuscenes_models/cbgs_voxel0075_res3d_centerpoint.yaml L6, L55-61
    "nus_0075": dict(kind="nus", pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], voxel_size=[0.075, 0.075, 0.2],
                     max_points=10, max_voxels={"train": 120000, "test": 160000}, num_features=5, config_id=3),
    # cfg #4: tools/cfgs/stage2_advmix/centerpoint_5_lab_95_unlab_nus_frames_advmix.yaml L13-14, L56-80
    "toda_stage2": dict(kind="nus", pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 4.8], voxel_size=[0.075, 0.075, 0.2],
                        max_points=10, max_voxels={"train": 120000, "test": 160000}, num_features=4, config_id=4,
                        sweeps=1),
}


def grid_size_xyz(pc_range, voxel_size):
    """round((hi-lo)/vsize), as data_processor.py L117-118."""
    r = np.asarray(pc_range, dtype=np.float64)
    return np.round((r[3:6] - r[0:3]) / np.asarray(voxel_size, dtype=np.float64)).astype(np.int64)


def _scene(rng, n_obj, r_min, r_max):
    """Random vertical structures: azimuth centre, half-angle, range, top height above ground.
    Mix: car-sized boxes, poles / trunks, and far building walls (area-uniform placement)."""
    az0 = rng.uniform(-np.pi, np.pi, n_obj)
    rng_m = np.sqrt(r_min ** 2 + (r_max ** 2 - r_min ** 2) * rng.random(n_obj))
    kind = rng.choice(3, n_obj, p=[0.6, 0.25, 0.15])
    width = np.where(kind == 0, rng.uniform(1.8, 4.5, n_obj),
                     np.where(kind == 1, rng.uniform(0.3, 1.0, n_obj), rng.uniform(8.0, 25.0, n_obj)))
    top = np.where(kind == 0, rng.uniform(1.4, 2.0, n_obj),
                   np.where(kind == 1, rng.uniform(4.0, 9.0, n_obj), rng.uniform(6.0, 15.0, n_obj)))
    rng_m = np.where(kind == 2, np.maximum(rng_m, 25.0), rng_m)
    half = np.arctan2(width * 0.5, rng_m)
    return az0, half, rng_m, top


def _cast(az, el, origin_xy, h_sensor, scene, max_range):
    """Distance along each ray (az, el) to the nearest of ground plane / structures; inf = no return."""
    az0, half, r_obj, top = scene
    cos_el, sin_el = np.cos(el), np.sin(el)
    with np.errstate(divide="ignore"):
        t_ground = np.where(sin_el < -1e-4, h_sensor / -sin_el, np.inf)
    best = t_ground
    # structures: hit if azimuth inside the interval and the ray is below the top at that range
    ox, oy = origin_xy
    for j in range(len(az0)):
        # object centre seen from the (shifted) sweep origin
        cx, cy = r_obj[j] * np.cos(az0[j]) - ox, r_obj[j] * np.sin(az0[j]) - oy
        rj = np.hypot(cx, cy)
        aj = np.arctan2(cy, cx)
        d = np.abs(np.angle(np.exp(1j * (az - aj))))
        t = rj / np.maximum(cos_el, 1e-3)
        z_at = h_sensor + t * sin_el  # height above ground where the ray meets the structure
        hit = (d < half[j]) & (z_at < top[j]) & (z_at > 0.0)
        best = np.where(hit & (t < best), t, best)
    return np.where(best <= max_range, best, np.inf)


def nuscenes_frame(rng, sweeps=10, num_features=5):
    """~266k points for 10 sweeps (32 beams x ~1090 azimuth steps x 10, 8% dropout, no-return rays removed)."""
    n_az, n_beam, h = 1090, 32, 1.84
    el_b = np.deg2rad(np.linspace(-30.67, 10.67, n_beam))
    scene = _scene(rng, 70, 4.0, 60.0)
    out = []
    speed = rng.uniform(2.0, 14.0)  # m/s ego motion along +x
    for s in range(sweeps):
        az = np.linspace(-np.pi, np.pi, n_az, endpoint=False) + rng.uniform(0, 2 * np.pi / n_az)
        A, E = np.meshgrid(az, el_b)
        A, E = A.ravel(), E.ravel()
        lag = 0.05 * s
        shift = -speed * lag  # past sweeps were taken from behind the current pose
        t = _cast(A, E, (shift, 0.0), h, scene, 70.0)
        keep = np.isfinite(t) & (rng.random(t.shape) > 0.08)
        t = t[keep] * (1.0 + 0.002 * rng.standard_normal(keep.sum()))
        a, e = A[keep], E[keep]
        x = t * np.cos(e) * np.cos(a) + shift
        y = t * np.cos(e) * np.sin(a)
        z = t * np.sin(e)  # sensor-centred: ground near z = -h
        ego = (np.abs(x) < 1.0) & (np.abs(y) < 1.0)
        inten = rng.random(t.shape)
        cols = [x, y, z, inten, np.full_like(x, lag)]
        out.append(np.stack(cols[:num_features], axis=1)[~ego])
    return np.concatenate(out).astype(np.float32)


def waymo_frame(rng, num_features=5):
    """~180k points: 64-beam top lidar (2650 azimuth steps) + four short-range side lidars; ground z~0."""
    h = 2.0
    scene = _scene(rng, 90, 5.0, 70.0)
    el_b = np.deg2rad(np.linspace(-17.6, 2.4, 64))
    az = np.linspace(-np.pi, np.pi, 2650, endpoint=False)
    A, E = np.meshgrid(az, el_b)
    A, E = A.ravel(), E.ravel()
    t = _cast(A, E, (0.0, 0.0), h, scene, 75.0)
    keep = np.isfinite(t) & (rng.random(t.shape) > 0.05)
    parts = [(t[keep], A[keep], E[keep], 0.0, 0.0, h)]
    for (sx, sy, a0) in [(4.0, 0.0, 0.0), (-1.0, 0.0, np.pi), (1.5, 1.0, np.pi / 2), (1.5, -1.0, -np.pi / 2)]:
        els = np.deg2rad(np.linspace(-60.0, 20.0, 40))
        azs = a0 + np.linspace(-np.pi / 2, np.pi / 2, 160)
        As, Es = np.meshgrid(azs, els)
        As, Es = As.ravel(), Es.ravel()
        ts = _cast(As, Es, (sx, sy), 0.8, scene, 20.0)
        ks = np.isfinite(ts) & (rng.random(ts.shape) > 0.05)
        parts.append((ts[ks], As[ks], Es[ks], sx, sy, 0.8))
    pts = []
    for (tt, a, e, sx, sy, hh) in parts:
        tt = tt * (1.0 + 0.002 * rng.standard_normal(tt.shape))
        x = tt * np.cos(e) * np.cos(a) + sx
        y = tt * np.cos(e) * np.sin(a) + sy
        z = hh + tt * np.sin(e)
        inten = np.tanh(rng.gamma(1.5, 0.4, tt.shape))
        elong = rng.random(tt.shape) * 0.3
        pts.append(np.stack([x, y, z, inten, elong][:num_features], axis=1))
    return np.concatenate(pts).astype(np.float32)


def mask_points_by_range(points, limit_range):
    """pcdet/utils/common_utils.py L60-63: inclusive on both ends, x/y only."""
    return (points[:, 0] >= limit_range[0]) & (points[:, 0] <= limit_range[3]) \
        & (points[:, 1] >= limit_range[1]) & (points[:, 1] <= limit_range[4])


def make_frame(config, frame_idx, shuffle=True, n_points=None):
    """One frame after the reference's host-side pre-steps (range mask L78-91, shuffle L93-103).
    Returns float32 (N, F) C-contiguous."""
    cfg = CONFIGS[config] if isinstance(config, str) else config
    rng = np.random.default_rng(1000 * cfg["config_id"] + frame_idx)
    if cfg["kind"] == "nus":
        pts = nuscenes_frame(rng, sweeps=cfg.get("sweeps", 10), num_features=cfg["num_features"])
    else:
        pts = waymo_frame(rng, num_features=cfg["num_features"])
    pts = pts[mask_points_by_range(pts, np.asarray(cfg["pc_range"], dtype=np.float32))]
    if n_points is not None:  # sweep config #5: resample to a requested size
        idx = rng.integers(0, pts.shape[0], n_points)
        jitter = (0.02 * rng.standard_normal((n_points, 3))).astype(np.float32)
        pts = pts[idx].copy()
        pts[:, :3] += jitter
        pts = pts[mask_points_by_range(pts, np.asarray(cfg["pc_range"], dtype=np.float32))]
    if shuffle:
        pts = pts[rng.permutation(pts.shape[0])]
    return np.ascontiguousarray(pts, dtype=np.float32)


def make_batch(config, batch_size, first_frame=0, shuffle=True):
    """List of per-frame point arrays plus the collated `points` tensor layout of
    pcdet/datasets/dataset.py L173-178: (sum N, 1+F) with the batch index in column 0."""
    frames = [make_frame(config, first_frame + i, shuffle) for i in range(batch_size)]
    cols = [np.pad(f, ((0, 0), (1, 0)), mode="constant", constant_values=i) for i, f in enumerate(frames)]
    return frames, np.concatenate(cols, axis=0)
