"""Seeded synthetic OPV2V-shaped scenes: no dataset, no network.

Per agent `a` of scene seed `s`: rng = default_rng(1000*s + a).  LiDAR = `rings` x `azimuth` rays from a
sensor at z=0 (lidar frame) onto a ground plane z=-1.9 and 30 yawed 4.5x2x1.6 m boxes inside +-90 m,
first hit, max range 120 m, N(0,0.02) noise, U(0,1) intensity, ego-vehicle points removed
(opencood/utils/pcd_utils.py:70-88) and one permutation (stands in for shuffle_points :91-95).
Poses: ego at the origin, others (x,y)~U(-40,40) m, yaw~U(-180,180) deg; pairwise_t_matrix follows
opencood/utils/transformation_utils.py:21-66 (x_to_world :264-308 + np.linalg.solve), identity padded to L.
"""
from __future__ import annotations

import numpy as np


def x_to_world(pose):
    x, y, z, roll, yaw, pitch = pose
    c_y, s_y = np.cos(np.radians(yaw)), np.sin(np.radians(yaw))
    c_r, s_r = np.cos(np.radians(roll)), np.sin(np.radians(roll))
    c_p, s_p = np.cos(np.radians(pitch)), np.sin(np.radians(pitch))
    m = np.identity(4)
    m[0, 3], m[1, 3], m[2, 3] = x, y, z
    m[0, 0] = c_p * c_y
    m[0, 1] = c_y * s_p * s_r - s_y * c_r
    m[0, 2] = -c_y * s_p * c_r - s_y * s_r
    m[1, 0] = s_y * c_p
    m[1, 1] = s_y * s_p * s_r + c_y * c_r
    m[1, 2] = -s_y * s_p * c_r + c_y * s_r
    m[2, 0] = s_p
    m[2, 1] = -c_p * s_r
    m[2, 2] = c_p * c_r
    return m


def pairwise_t_matrix(poses, max_cav):
    t = np.tile(np.eye(4), (max_cav, max_cav, 1, 1))
    tl = [x_to_world(p) for p in poses]
    for i in range(len(tl)):
        for j in range(len(tl)):
            if i != j:
                t[i, j] = np.linalg.solve(tl[j], tl[i])
    return t


def lidar_cloud(rng: np.random.Generator, rings=64, azimuth=1024, n_boxes=30):
    elev = np.radians(np.linspace(-25.0, 2.0, rings))
    azim = np.linspace(-np.pi, np.pi, azimuth, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing="ij")
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1).reshape(-1, 3)
    n = d.shape[0]
    t_hit = np.full(n, np.inf)
    # ground plane z = -1.9
    down = d[:, 2] < -1e-6
    t_hit[down] = -1.9 / d[down, 2]
    # boxes
    ctr = np.stack([rng.uniform(-90, 90, n_boxes), rng.uniform(-90, 90, n_boxes), np.full(n_boxes, -1.9 + 0.8)], 1)
    yaw = rng.uniform(-np.pi, np.pi, n_boxes)
    half = np.array([2.25, 1.0, 0.8])
    for b in range(n_boxes):
        c, s = np.cos(yaw[b]), np.sin(yaw[b])
        R = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])  # world -> box
        o = R @ (-ctr[b])
        dd = d @ R.T
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-half - o) / dd
            t2 = (half - o) / dd
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tmax >= tmin) & (tmax > 0) & (tmin > 0)
        t_hit = np.where(hit & (tmin < t_hit), tmin, t_hit)
    keep = np.isfinite(t_hit) & (t_hit <= 120.0)
    xyz = d[keep] * t_hit[keep, None] + rng.normal(0, 0.02, (int(keep.sum()), 3))
    inten = rng.uniform(0, 1, (xyz.shape[0], 1))
    pts = np.concatenate([xyz, inten], 1).astype(np.float32)
    ego = (pts[:, 0] >= -1.95) & (pts[:, 0] <= 2.95) & (pts[:, 1] >= -1.1) & (pts[:, 1] <= 1.1)
    pts = pts[~ego]
    return pts[rng.permutation(pts.shape[0])]


def scene(seed: int, n_agents: int = 5, max_cav: int = 5, rings=64, azimuth=1024):
    """Returns dict(points=[(P_a,4) f32], poses (n,6), pairwise_t_matrix (1,L,L,4,4) f64, record_len)."""
    pts, poses = [], []
    for a in range(n_agents):
        rng = np.random.default_rng(1000 * seed + a)
        pts.append(lidar_cloud(rng, rings, azimuth))
        if a == 0:
            poses.append([0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        else:
            poses.append([rng.uniform(-40, 40), rng.uniform(-40, 40), 0.0, 0.0, rng.uniform(-180, 180), 0.0])
    L = max(max_cav, n_agents)
    return {
        "points": pts,
        "poses": np.asarray(poses),
        "pairwise_t_matrix": pairwise_t_matrix(poses, L)[None],
        "record_len": np.asarray([n_agents], dtype=np.int64),
    }


def camera_rig(n_agents: int, n_cams: int = 4, H: int = 256, W: int = 704):
    """4 cameras yaw 0/100/-100/180 deg, fx=fy=560 px scaled to the 704x256 image; post_rots=I, post_trans=0.
    Returns rots (n,4,3,3), trans (n,4,3), intrins (n,4,3,3), post_rots, post_trans (f32 numpy)."""
    yaws = np.radians([0.0, 100.0, -100.0, 180.0])[:n_cams]
    fx = 560.0 * W / 800.0
    K = np.array([[fx, 0, W / 2], [0, fx, H / 2], [0, 0, 1.0]])
    # camera frame (x right, y down, z forward) -> ego (x forward, y left, z up)
    c2e = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    rots, trans = [], []
    for y in yaws:
        Rz = np.array([[np.cos(y), -np.sin(y), 0], [np.sin(y), np.cos(y), 0], [0, 0, 1.0]])
        rots.append(Rz @ c2e)
        trans.append(Rz @ np.array([1.5, 0.0, 0.0]) + np.array([0, 0, 1.6]))
    rots = np.tile(np.asarray(rots)[None], (n_agents, 1, 1, 1)).astype(np.float32)
    trans = np.tile(np.asarray(trans)[None], (n_agents, 1, 1)).astype(np.float32)
    intr = np.tile(K[None, None], (n_agents, n_cams, 1, 1)).astype(np.float32)
    post_rots = np.tile(np.eye(3)[None, None], (n_agents, n_cams, 1, 1)).astype(np.float32)
    post_trans = np.zeros((n_agents, n_cams, 3), dtype=np.float32)
    return rots, trans, intr, post_rots, post_trans
