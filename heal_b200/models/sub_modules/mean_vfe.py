"""MeanVFE (reference: opencood/models/sub_modules/mean_vfe.py:13-33): per-voxel mean of the point slots, the VFE of the SECOND encoder.

One launch of `heal_mean_vfe` (csrc/voxelize.cu): thread per (voxel, feature), sum over the T zero-padded slots divided by
max(num_points, 1) in the reference's order of operations; parameter-free, so there is no state dict to mirror -- only the
constructor signature, `get_output_feature_dim` and the batch_dict keys ('voxel_features' in: (M, T, C); out: (M, C))."""
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
