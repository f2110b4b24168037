"""The five BASELINE.json configs as `hypes['model']['args']` dictionaries at the reference's own yaml values
(SURVEY.md 8d "Configs restated").  Pure data: imported by bench.py (both arms), the -m gpu full-size parity tests and
the reference runner; imports neither heal_b200 nor oracle.

yaml sources (relative to the reference root, opencood/hypes_yaml/opv2v/MoreModality/):
  C1  HEAL/stage1/m1_pyramid.yaml:100-109 encoder block inside models/point_pillar.py + base_bev_backbone [3,5,8]
  C2  HEAL/stage1/m1_pyramid.yaml:91-138
  C3  3_modality_end2end_training/m1m2m3_attfuse.yaml:198-233 (m3 block), voxel 0.1^3 T=5 (:69-83)
  C4  HEAL/final_infer/m1m2m3m4.yaml:52-57,188-189 (m2 lift_splat_shoot, Resnet101 trunk) + m1
  C5  C2 with max_cav 8
"""
import copy

import numpy as np

RANGE = [-102.4, -102.4, -3, 102.4, 102.4, 1]          # inference.py:34 / lidar_attfuse.yaml:17
CAM_RANGE = [-51.2, -51.2, -3, 51.2, 51.2, 1]
PILLAR_VOXEL = [0.4, 0.4, 4]
SECOND_VOXEL = [0.1, 0.1, 0.1]
DIR_ARGS = {"dir_offset": 0.7853, "num_bins": 2, "anchor_yaw": [0, 90]}
SHRINK = {"kernal_size": [3], "stride": [1], "padding": [1], "dim": [256], "input_dim": 384}


def pillar_encoder_args(rng=RANGE):
    return {"voxel_size": list(PILLAR_VOXEL), "lidar_range": list(rng),
            "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
            "point_pillar_scatter": {"num_features": 64}}


def second_encoder_args(rng=RANGE):
    return {"voxel_size": list(SECOND_VOXEL), "lidar_range": list(rng), "mean_vfe": {"num_point_features": 4},
            "spconv": {"num_features_in": 4, "num_features_out": 64}, "map2bev": {"feature_num": 128}}


def lss_encoder_args(final_dim=(256, 704)):
    return {"grid_conf": {"xbound": [-51.2, 51.2, 0.4], "ybound": [-51.2, 51.2, 0.4], "zbound": [-10, 10, 20.0],
                          "ddiscr": [2, 50, 48], "mode": "LID"},
            "data_aug_conf": {"final_dim": list(final_dim)}, "img_downsample": 8, "img_features": 128,
            "camera_encoder": "Resnet101", "use_depth_gt": False, "depth_supervision": False}


def pyramid_backbone_args():
    return {"resnext": True, "layer_nums": [3, 5, 8], "layer_strides": [1, 2, 2], "num_filters": [64, 128, 256],
            "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128], "anchor_number": 2}


def c1_args(rng=RANGE):
    """models/point_pillar.py single agent."""
    a = pillar_encoder_args(rng)
    # yaml_utils.load_point_pillar_params (hypes_yaml/yaml_utils.py:121-135) writes the grid size into the scatter args
    a["point_pillar_scatter"]["grid_size"] = np.round((np.array(rng[3:6]) - np.array(rng[0:3])) / np.array(PILLAR_VOXEL)).astype(np.int64)
    a.update({"anchor_number": 2,
              "base_bev_backbone": {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
                                    "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]},
              "shrink_header": copy.deepcopy(SHRINK), "dir_args": copy.deepcopy(DIR_ARGS)})
    return a


def c2_args(rng=RANGE):
    """heter_pyramid_collab, m1 = PointPillars, PyramidFusion ResNeXt."""
    return {"lidar_range": list(rng), "supervise_single": True,
            "m1": {"core_method": "point_pillar", "sensor_type": "lidar", "encoder_args": pillar_encoder_args(rng),
                   "backbone_args": {"layer_nums": [3], "layer_strides": [2], "num_filters": [64]},
                   "aligner_args": {"core_method": "identity"}},
            "fusion_backbone": pyramid_backbone_args(), "shrink_header": copy.deepcopy(SHRINK),
            "in_head": 256, "anchor_number": 2, "dir_args": copy.deepcopy(DIR_ARGS)}


def c3_args(rng=RANGE):
    """heter_model_baseline with the SECOND block (registered as m1, the only modality) + AttFusion."""
    return {"lidar_range": list(rng), "ego_modality": "m1",
            "m1": {"core_method": "second", "sensor_type": "lidar", "encoder_args": second_encoder_args(rng),
                   "backbone_args": {"layer_nums": [3, 5, 8], "layer_strides": [1, 2, 2], "num_filters": [64, 128, 256],
                                     "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128], "inplanes": 128},
                   "shrink_header": copy.deepcopy(SHRINK)},
            "fusion_method": "att", "att": {"feat_dim": 256},
            "in_head": 256, "anchor_number": 2, "dir_args": copy.deepcopy(DIR_ARGS)}


def c4_args(rng=RANGE, final_dim=(256, 704)):
    """heter_pyramid_collab with m1 (PointPillars) + m2 (Lift-Splat-Shoot, Resnet101 trunk); agents [m1, m2, m2]."""
    a = c2_args(rng)
    enc = lss_encoder_args(final_dim)
    a["m2"] = {"core_method": "lift_splat_shoot", "sensor_type": "camera", "encoder_args": enc,
               "camera_mask_args": {"grid_conf": copy.deepcopy(enc["grid_conf"])},
               "backbone_args": {"layer_nums": [3], "layer_strides": [2], "num_filters": [64], "inplanes": 128},
               "aligner_args": {"core_method": "identity"}}
    return a


def c5_args(rng=RANGE):
    return c2_args(rng)


WORKLOADS = {
    "c1": {"title": "configs[0]: point_pillar single agent, 20k-point synthetic cloud, range +-102.4 m, 512x512 pillars",
           "agents": 1, "rings": 20, "azimuth": 1000},
    "c2": {"title": "configs[1]: heter_pyramid_collab (PointPillars m1 + PyramidFusion ResNeXt), 5 agents x 64-line LiDAR "
                    "(~58k pts/agent), range +-102.4 m, 512x512 pillars @0.4 m, fusion map 256x256, batch 1 scene",
           "agents": 5, "rings": 64, "azimuth": 1024},
    "c3": {"title": "configs[2]: heter_model_baseline SECOND (VoxelBackBone8x sparse conv, 0.1 m voxels, 2048x2048x40 grid) + "
                    "BaseBEVBackbone + AttFusion, 5 agents x 64-line LiDAR, BEV map 256x256",
           "agents": 5, "rings": 64, "azimuth": 1024},
    "c4": {"title": "configs[3]: heter_pyramid_collab hetero: agents [PointPillars, LSS 4-cam 704x256, LSS], bf16, fusion map 256x256",
           "agents": 3, "rings": 64, "azimuth": 1024},
    "c5": {"title": "configs[4]: 8-agent scene, heter_pyramid_collab (C2 model, max_cav 8), one agent per GPU + NCCL BEV all-gather",
           "agents": 8, "rings": 64, "azimuth": 1024},
}
