"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU voxel generator (see voxelizer.c for provenance; PARITY
UNPINNED against real spconv) + the reference's collate (sp_voxel_preprocessor.py:145-174)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_voxelizer.so")


def build_c(force=False):
    src = os.path.join(_HERE, "voxelizer.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off",
                               src, "-o", _SO, "-lm"])
    return _SO


def grid_size_of(lidar_range, voxel_size):
    g = (np.array(lidar_range[3:6]) - np.array(lidar_range[0:3])) / np.array(voxel_size)
    return np.round(g).astype(np.int64)


def points_to_voxel_c(points, voxel_size, lidar_range, max_points, max_voxels):
    """C restatement (fast; used for full-size clouds and as bench.py's cpu_baseline voxelizer)."""
    lib = ctypes.CDLL(build_c())
    pts = np.ascontiguousarray(points, dtype=np.float32)
    n, f = pts.shape
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(lidar_range, dtype=np.float32)
    grid = grid_size_of(lidar_range, voxel_size).astype(np.int32)
    voxels = np.empty((max_voxels, max_points, f), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    npv = np.empty((max_voxels,), dtype=np.int32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.oracle_points_to_voxel.restype = ctypes.c_int
    m = lib.oracle_points_to_voxel(fp(pts), n, f, fp(vs), fp(rg), fp(grid), int(max_points), int(max_voxels),
                                   fp(voxels), fp(coors), fp(npv))
    assert m >= 0
    return voxels[:m].copy(), coors[:m].copy(), npv[:m].copy()


def points_to_voxel_py(points, voxel_size, lidar_range, max_points, max_voxels):
    """Pure-Python restatement of the same loop (small cases only; cross-checks the C build)."""
    pts = np.asarray(points, dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    rg = np.asarray(lidar_range, dtype=np.float32)
    grid = grid_size_of(lidar_range, voxel_size)
    table = {}
    voxels, coors, npv = [], [], []
    for i in range(pts.shape[0]):
        c = [0, 0, 0]
        ok = True
        for j in range(3):
            q = np.float32(np.float32(pts[i, j] - rg[j]) / vs[j])
            fl = np.floor(q)
            if not (fl >= 0 and fl < grid[j]):
                ok = False
                break
            c[2 - j] = int(fl)
        if not ok:
            continue
        key = tuple(c)
        idx = table.get(key, -1)
        if idx == -1:
            if len(voxels) >= max_voxels:
                continue
            idx = len(voxels)
            table[key] = idx
            voxels.append(np.zeros((max_points, pts.shape[1]), dtype=np.float32))
            coors.append(c)
            npv.append(0)
        if npv[idx] < max_points:
            voxels[idx][npv[idx]] = pts[i]
            npv[idx] += 1
    if not voxels:
        return (np.zeros((0, max_points, pts.shape[1]), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    return np.stack(voxels), np.asarray(coors, dtype=np.int32), np.asarray(npv, dtype=np.int32)


def collate(per_agent):
    """sp_voxel_preprocessor.py:145-174: concat agents, prepend the agent index column."""
    vf = np.concatenate([a[0] for a in per_agent])
    npv = np.concatenate([a[2] for a in per_agent])
    coords = np.concatenate([np.pad(a[1], ((0, 0), (1, 0)), mode="constant", constant_values=i)
                             for i, a in enumerate(per_agent)])
    return {"voxel_features": vf, "voxel_coords": coords.astype(np.int32), "voxel_num_points": npv}
