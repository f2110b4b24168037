"""Registry mirroring opencood/data_utils/pre_processor/__init__.py:10-32 for the GPU path."""
from .gpu_voxel_preprocessor import GpuVoxelPreprocessor

__all__ = {'GpuVoxelPreprocessor': GpuVoxelPreprocessor}


def build_preprocessor(preprocess_cfg, train):
    name = preprocess_cfg['core_method']
    if name not in __all__:
        raise KeyError(f"{name}: heal_b200 provides {sorted(__all__)}; the CPU pre-processors stay in opencood")
    return __all__[name](preprocess_params=preprocess_cfg, train=train)
