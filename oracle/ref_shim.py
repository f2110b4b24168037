"""Import shim for the UNMODIFIED reference (yifanlu0227/HEAL).

TEST INFRASTRUCTURE ONLY.  The reference package is imported from `oracle/_ref/` (the byte-for-byte copy made by
`oracle/build_ref.py`; git-ignored, travels to the GPU box) or, in the build container, straight from /root/reference.
Users: oracle/make_golden.py (golden vectors), oracle/ref_runner.py (bench.py's reference arm and baseline legs) and the
boundary tests.  Nothing under heal_b200/ imports this file.

The reference imports a handful of cosmetic / absent third-party packages at module top
(SURVEY.md §8c).  They are replaced by MagicMock stubs; spconv is absent, so every spconv-backed
class is unusable here (its restatement lives in oracle/voxelizer.* and oracle/sparse_conv.py).
"""
import os
import sys
import types
from unittest.mock import MagicMock

_LOCAL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("HEAL_REFERENCE_ROOT") or (_LOCAL if os.path.isdir(os.path.join(_LOCAL, "opencood")) else "/root/reference")

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
    # heal_b200.install.install_into_opencood() aliases product modules under opencood.* names; the reference side must never
    # resolve to them, so such aliases are dropped here and re-imported from the reference tree on demand
    for k in [k for k, v in sys.modules.items() if k.startswith("opencood") and getattr(v, "__name__", k).startswith("heal_b200")]:
        del sys.modules[k]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return REFERENCE_ROOT
