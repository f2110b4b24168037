"""CPU restatement of the reference's host-side point filters (TEST INFRASTRUCTURE ONLY; numpy).

Follows opencood/utils/pcd_utils.py: mask_points_by_range :41-67 (strict comparisons on x, y, z), mask_ego_points :70-88
(drop -1.95 <= x <= 2.95 and -1.1 <= y <= 1.1), shuffle_points :91-95 (rows re-ordered by a permutation), in the order
intermediate_heter_fusion_dataset.py:141-173 applies them (shuffle -> ego mask -> range mask).  Pinned against the unmodified
reference functions by tests/test_oracle_golden.py::test_point_filters_match_reference when oracle/_ref (or /root/reference) exists.
"""
import numpy as np


def mask_points_by_range(points, limit_range):
    m = ((points[:, 0] > limit_range[0]) & (points[:, 0] < limit_range[3]) & (points[:, 1] > limit_range[1])
         & (points[:, 1] < limit_range[4]) & (points[:, 2] > limit_range[2]) & (points[:, 2] < limit_range[5]))
    return points[m]


def mask_ego_points(points):
    m = (points[:, 0] >= -1.95) & (points[:, 0] <= 2.95) & (points[:, 1] >= -1.1) & (points[:, 1] <= 1.1)
    return points[np.logical_not(m)]


def filter_cloud(points, limit_range, perm=None, remove_ego=True):
    """shuffle (given permutation) -> mask_ego_points -> mask_points_by_range for one agent's (n,4) cloud."""
    if perm is not None:
        points = points[perm]
    if remove_ego:
        points = mask_ego_points(points)
    return mask_points_by_range(points, limit_range)
