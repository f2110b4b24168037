"""Import shim for the UNMODIFIED reference (yifanlu0227/HEAL) mounted at /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py, in the build container, to run the
reference's own PyTorch modules on CPU and dump golden vectors under tests/golden/.  Nothing under
heal_b200/ imports this file, and nothing that runs on the GPU box does (/root/reference does not
exist there).

The reference imports a handful of cosmetic / absent third-party packages at module top
(SURVEY.md §8c).  They are replaced by MagicMock stubs; spconv is absent, so every spconv-backed
class is unusable here (its restatement lives in oracle/voxelizer.* and oracle/sparse_conv.py).
"""
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("HEAL_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm", "icecream", "termcolor",
    "shapely", "shapely.geometry", "open3d", "pyquaternion", "efficientnet_pytorch",
    "tensorboardX", "timm", "timm.models", "timm.models.layers", "timm.models.registry",
    "spconv", "spconv.pytorch", "spconv.utils", "cumm", "torch_scatter", "h5py", "easydict",
    "skimage", "skimage.transform", "cv2",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "opencood"))


def install():
    """Put the reference on sys.path with stubbed cosmetic deps. Idempotent."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except Exception:
            m = MagicMock(name=name)
            m.__path__ = []  # behave as a package
            m.__spec__ = None
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return REFERENCE_ROOT
