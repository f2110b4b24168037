"""Registry mirroring opencood/data_utils/post_processor/__init__.py for the GPU path."""
from .voxel_postprocessor import VoxelPostprocessor

__all__ = {'VoxelPostprocessor': VoxelPostprocessor}


def build_postprocessor(anchor_cfg, train):
    name = anchor_cfg['core_method']
    if name not in __all__:
        raise KeyError(f"{name}: heal_b200 provides {sorted(__all__)}")
    return __all__[name](anchor_params=anchor_cfg, train=train)
