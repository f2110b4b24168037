"""HeightCompression mirror (opencood/models/sub_modules/height_compression.py:10-27) on heal_sparse_to_bev."""
import torch.nn as nn

from ... import ops


class HeightCompression(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg['feature_num']

    def forward_act(self, sparse_tensor):
        return ops.sparse_to_bev(sparse_tensor)          # Act f32 (N, H, W, C*D)

    def forward(self, batch_dict):
        bev = self.forward_act(batch_dict['encoded_spconv_tensor'])
        batch_dict['spatial_features'] = ops.act_to_nchw(bev)
        batch_dict['spatial_features_stride'] = batch_dict['encoded_spconv_tensor_stride']
        return batch_dict
