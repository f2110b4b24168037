"""MeanVFE mirror (opencood/models/sub_modules/mean_vfe.py:13-33) on heal_mean_vfe."""
import torch.nn as nn

from ... import ops


class MeanVFE(nn.Module):
    def __init__(self, model_cfg, num_point_features, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, batch_dict, **kwargs):
        batch_dict['voxel_features'] = ops.mean_vfe(batch_dict['voxel_features'].contiguous(),
                                                    batch_dict['voxel_num_points'])
        return batch_dict
