"""Kept for import compatibility: the seeded scene generator lives in `workloads/synth.py` (shared by the product's
bench/tests and by the reference arm, which must not import this package)."""
from workloads.synth import *  # noqa: F401,F403
from workloads.synth import x_to_world, pairwise_t_matrix, lidar_cloud, scene, camera_rig  # noqa: F401
