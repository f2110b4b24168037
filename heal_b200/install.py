"""Drop the heal_b200 mirrors into an importable `opencood` tree so that the reference's own entry points
(opencood/tools/train_utils.create_model -> importlib `opencood.models.<core_method>`, heter_pyramid_collab's
`opencood.models.heter_encoders` lookup) resolve to the B200 implementations.  See INTEGRATION.md."""
import importlib
import sys

_MAP = {
    "opencood.models.heter_pyramid_collab": "heal_b200.models.heter_pyramid_collab",
    "opencood.models.heter_model_baseline": "heal_b200.models.heter_model_baseline",
    "opencood.models.point_pillar": "heal_b200.models.point_pillar",
    "opencood.models.heter_encoders": "heal_b200.models.heter_encoders",
    "opencood.models.fuse_modules.pyramid_fuse": "heal_b200.models.fuse_modules.pyramid_fuse",
    "opencood.models.sub_modules.pillar_vfe": "heal_b200.models.sub_modules.pillar_vfe",
    "opencood.models.sub_modules.point_pillar_scatter": "heal_b200.models.sub_modules.point_pillar_scatter",
    "opencood.models.sub_modules.mean_vfe": "heal_b200.models.sub_modules.mean_vfe",
    "opencood.models.sub_modules.sparse_backbone_3d": "heal_b200.models.sub_modules.sparse_backbone_3d",
    "opencood.models.sub_modules.height_compression": "heal_b200.models.sub_modules.height_compression",
    "opencood.models.sub_modules.base_bev_backbone": "heal_b200.models.sub_modules.base_bev_backbone",
    "opencood.models.sub_modules.base_bev_backbone_resnet": "heal_b200.models.sub_modules.base_bev_backbone_resnet",
    "opencood.models.sub_modules.resblock": "heal_b200.models.sub_modules.resblock",
    "opencood.models.sub_modules.downsample_conv": "heal_b200.models.sub_modules.downsample_conv",
}


def install_gpu_postprocessor():
    """Opt-in: make the reference's post-processor registry (opencood/data_utils/post_processor/__init__.py `__all__`) hand out a
    SUBCLASS of its own VoxelPostprocessor whose `post_process` runs on the GPU for early / intermediate fusion and falls back to
    the reference implementation for everything else (late fusion, iou_preds, CPU tensors).  All label / collate / gt-box methods
    the datasets call are inherited unchanged."""
    from .data_utils.post_processor.voxel_postprocessor import make_reference_subclass
    reg = importlib.import_module("opencood.data_utils.post_processor")
    ref_cls = importlib.import_module("opencood.data_utils.post_processor.voxel_postprocessor").VoxelPostprocessor
    cls = make_reference_subclass(ref_cls)
    reg.__all__["VoxelPostprocessor"] = cls
    return cls


def install_into_opencood(only=None):
    """Alias the mirrors under their opencood module names (call BEFORE opencood.tools imports the models)."""
    done = []
    for ref_name, ours in _MAP.items():
        if only is not None and ref_name not in only:
            continue
        sys.modules[ref_name] = importlib.import_module(ours)
        done.append(ref_name)
    return done
