"""Moved to `workloads/procedural.py` (a deterministic parameter table is workload data, not oracle compute)."""
from workloads.procedural import *  # noqa: F401,F403
from workloads.procedural import make_tensor, make_state_dict, shapes_of  # noqa: F401
