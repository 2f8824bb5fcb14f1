"""Deterministic synthetic LiDAR-shaped clouds (our own generator, frozen).

KITTI-shaped: 64 beams, ~130k raw hits -> ~16-18k points per cloud after 0.3 m voxelisation.
nuScenes-shaped: 32 beams, ~32k raw hits -> ~9k points per cloud at 0.3 m.
A "pair" = two scans of the same scene (same `seed`) from poses pose_x = 0 and pose_x = d.
"""
import numpy as np


def lidar(seed, beams, el_lo, el_hi, n_az, h, rmax, n_cyl=80, wall_lo=20., wall_hi=35., wall_h=10., pose_x=0.0):
    rng = np.random.default_rng(seed)                       # scene RNG: identical for both clouds of a pair
    el = np.deg2rad(np.linspace(el_lo, el_hi, beams))[:, None]
    az = np.linspace(0, 2 * np.pi, n_az, endpoint=False)[None, :]
    dx, dy, dz = np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)
    r = np.full(dx.shape, np.inf)
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.minimum(r, np.where(dz < 0, -h / dz, np.inf))                    # ground plane z = -h
        for sgn in (+1, -1):                                                    # two street walls
            W = rng.uniform(wall_lo, wall_hi)
            rw = np.where(sgn * dy > 0, W / (sgn * dy), np.inf)
            xh = np.where(np.isfinite(rw), pose_x + rw * dx, 0.0)
            gap = (np.floor(xh / 12.0).astype(np.int64) % 4 == 0)               # 12 m opening every 48 m
            zh = rw * dz
            r = np.minimum(r, np.where((zh < wall_h - h) & (zh > -h) & ~gap, rw, np.inf))
        cyl = np.stack([rng.uniform(-60, 60, n_cyl), rng.uniform(-wall_hi, wall_hi, n_cyl),
                        rng.uniform(0.3, 2.0, n_cyl), rng.uniform(1.0, 4.0, n_cyl)], 1)
        for cx, cy, rad, top in cyl:                                            # vertical cylinders
            cx = cx - pose_x
            a = dx * dx + dy * dy; b = -2 * (dx * cx + dy * cy); c = cx * cx + cy * cy - rad * rad
            disc = b * b - 4 * a * c
            t = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * a)
            r = np.minimum(r, np.where((disc > 0) & (t > 0) & (t * dz < top - h) & (t * dz > -h), t, np.inf))
    m = np.isfinite(r) & (r < rmax) & (r > 2.0)
    r = r + np.random.default_rng(seed * 7919 + int(pose_x * 1000) + 1).normal(0, 0.02, r.shape)
    with np.errstate(invalid='ignore'):
        pts = np.stack([r * dx, r * dy, r * dz], -1)
    return pts[m].astype(np.float32)


def kitti(seed, pose_x=0.0):
    return lidar(seed, 64, -24.8, 2.0, 2083, 1.73, 120.0, pose_x=pose_x)


def nusc(seed, pose_x=0.0):
    return lidar(seed, 32, -30.0, 10.0, 1090, 1.84, 70.0, pose_x=pose_x)


def pair_pose(seed, distant=False):
    """Sensor displacement (metres along world x) between the two scans of pair `seed`."""
    return float(np.random.default_rng(10_000 + seed).uniform(5.0, 50.0 if distant else 20.0))


def pair_raw(seed, kind="kitti", distant=False):
    """Two raw scans of scene `seed`. Ordinary pair: second pose U(5,20) m; distant (LoKITTI-like): U(5,50) m.
    Both clouds are in their own sensor frame: world = sensor + (pose_x, 0, 0)."""
    gen = kitti if kind == "kitti" else nusc
    d = pair_pose(seed, distant)
    return gen(seed, 0.0), gen(seed, d)


def small_cloud(seed, n=2000, extent=(20.0, 12.0, 3.0)):
    """Small generic test cloud: points on a few noisy planes + uniform clutter (fast to generate, seconds-size oracle)."""
    rng = np.random.default_rng(seed)
    n1 = n // 2
    ground = np.stack([rng.uniform(-extent[0], extent[0], n1), rng.uniform(-extent[1], extent[1], n1),
                       rng.normal(0, 0.03, n1)], 1)
    n2 = n // 4
    wall = np.stack([rng.uniform(-extent[0], extent[0], n2), np.full(n2, extent[1] * 0.7) + rng.normal(0, 0.03, n2),
                     rng.uniform(0, extent[2], n2)], 1)
    n3 = n - n1 - n2
    clutter = np.stack([rng.uniform(-extent[0], extent[0], n3), rng.uniform(-extent[1], extent[1], n3),
                        rng.uniform(0, extent[2], n3)], 1)
    return np.concatenate([ground, wall, clutter], 0).astype(np.float32)
