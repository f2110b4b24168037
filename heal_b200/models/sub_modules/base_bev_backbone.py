"""BaseBEVBackbone mirror (opencood/models/sub_modules/base_bev_backbone.py:7-156)."""
import numpy as np
import torch
import torch.nn as nn

from ... import ops
from ...engine import conv_bn_act, require_eval
from .base_bev_backbone_resnet import decode_levels


class BaseBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = model_cfg.get('layer_nums', [])
        layer_strides = model_cfg.get('layer_strides', [])
        num_filters = model_cfg.get('num_filters', [])
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        upsample_strides = model_cfg.get('upsample_strides', [])
        num_upsample_filters = model_cfg.get('num_upsample_filter', [])
        assert len(upsample_strides) == len(num_upsample_filters)
        num_levels = len(layer_nums)
        self.num_levels = num_levels
        c_in_list = [input_channels, *num_filters[:-1]]
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        for idx in range(num_levels):
            cur = [nn.ZeroPad2d(1),
                   nn.Conv2d(c_in_list[idx], num_filters[idx], 3, stride=layer_strides[idx], padding=0, bias=False),
                   nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(layer_nums[idx]):
                cur.extend([nn.Conv2d(num_filters[idx], num_filters[idx], 3, padding=1, bias=False),
                            nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()])
            self.blocks.append(nn.Sequential(*cur))
            if len(upsample_strides) > 0:
                stride = upsample_strides[idx]
                if stride >= 1:
                    self.deblocks.append(nn.Sequential(
                        nn.ConvTranspose2d(num_filters[idx], num_upsample_filters[idx], stride, stride=stride, bias=False),
                        nn.BatchNorm2d(num_upsample_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()))
                else:
                    k = int(np.round(1 / stride))
                    self.deblocks.append(nn.Sequential(
                        nn.Conv2d(num_filters[idx], num_upsample_filters[idx], k, stride=k, bias=False),
                        nn.BatchNorm2d(num_upsample_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()))
        if len(upsample_strides) > num_levels:
            raise NotImplementedError("extra trailing deblock is not used by any HEAL yaml")
        self.num_bev_features = sum(num_upsample_filters)

    def multiscale_nhwc(self, x):
        feats = []
        for blk in self.blocks:
            mods = list(blk)
            x = conv_bn_act(x, mods[1], mods[2], relu=True, extra_pad=1)   # ZeroPad2d(1) + conv(pad 0)
            i = 4
            while i < len(mods):
                x = conv_bn_act(x, mods[i], mods[i + 1], relu=True)
                i += 3
            feats.append(x)
        return feats

    def decode_nhwc(self, feats):
        return decode_levels(self.deblocks, feats)

    def forward(self, data_dict):
        require_eval(self)
        feats = self.multiscale_nhwc(ops.to_act(data_dict['spatial_features']))
        data_dict['spatial_features_2d'] = ops.act_to_nchw(self.decode_nhwc(feats))
        return data_dict

    def get_multiscale_feature(self, spatial_features):
        return [ops.act_to_nchw(f) for f in self.multiscale_nhwc(ops.to_act(spatial_features))]

    def decode_multiscale_feature(self, x):
        return ops.act_to_nchw(self.decode_nhwc([ops.to_act(f) for f in x]))
