"""ResNetBEVBackbone mirror (opencood/models/sub_modules/base_bev_backbone_resnet.py:13-142)."""
import numpy as np
import torch
import torch.nn as nn

from ... import ops
from ...engine import conv_bn_act, require_eval, act_fmt
from .resblock import ResNetModified, BasicBlock


def decode_levels(deblocks, feats):
    """deblocks (ConvTranspose2d k==s | Conv2d k==s, + BN + ReLU) with the channel concat written in place
    into one (N,H,W,sum C) buffer (base_bev_backbone_resnet.py:127-142 / base_bev_backbone.py:139-156)."""
    if len(deblocks) == 0:
        if len(feats) == 1:
            return feats[0]
        raise NotImplementedError("multi-level backbone without deblocks is not used by any HEAL yaml")
    couts = [d[0].out_channels for d in deblocks]
    d0 = deblocks[0][0]
    if isinstance(d0, nn.ConvTranspose2d):
        H0, W0 = feats[0].H * d0.stride[0], feats[0].W * d0.stride[0]
    else:
        H0, W0 = feats[0].H // d0.stride[0], feats[0].W // d0.stride[0]
    out = ops.act_empty(feats[0].N, H0, W0, sum(couts), act_fmt(), feats[0].device)
    off = 0
    for f, d, c in zip(feats, deblocks, couts):
        conv_bn_act(f, d[0], d[1], relu=True, out=out, out_coffset=off)
        off += c
    return out


class ResNetBEVBackbone(nn.Module):
    def __init__(self, model_cfg, input_channels=64):
        super().__init__()
        self.model_cfg = model_cfg
        layer_nums = model_cfg.get('layer_nums', [])
        layer_strides = model_cfg.get('layer_strides', [])
        num_filters = model_cfg.get('num_filters', [])
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        upsample_strides = model_cfg.get('upsample_strides', [])
        num_upsample_filters = model_cfg.get('num_upsample_filter', [])
        assert len(upsample_strides) == len(num_upsample_filters)
        self.resnet = ResNetModified(BasicBlock, layer_nums, layer_strides, num_filters,
                                     inplanes=model_cfg.get('inplanes', 64))
        self.num_levels = len(layer_nums)
        self.deblocks = nn.ModuleList()
        for idx in range(self.num_levels):
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
        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > self.num_levels:
            raise NotImplementedError("extra trailing deblock (upsample_strides longer than levels) is not used by any HEAL yaml")
        self.num_bev_features = c_in

    # ---- channels-last internals (Act in / Act out) ---------------------------------------------
    def multiscale_nhwc(self, x, outs=None):
        return self.resnet.forward_nhwc(x, outs=outs)

    def decode_nhwc(self, feats):
        return decode_levels(self.deblocks, feats)

    # ---- reference API --------------------------------------------------------------------------
    def forward(self, data_dict):
        require_eval(self)
        feats = self.multiscale_nhwc(ops.to_act(data_dict['spatial_features']))
        data_dict['spatial_features_2d'] = ops.act_to_nchw(self.decode_nhwc(feats))
        return data_dict

    def get_multiscale_feature(self, spatial_features):
        require_eval(self)
        return [ops.act_to_nchw(f) for f in self.multiscale_nhwc(ops.to_act(spatial_features))]

    def decode_multiscale_feature(self, x):
        require_eval(self)
        return ops.act_to_nchw(self.decode_nhwc([ops.to_act(f) for f in x]))

    def get_layer_i_feature(self, spatial_features, layer_i):
        x = ops.to_act(spatial_features)
        for blk in getattr(self.resnet, f"layer{layer_i}"):
            x = blk.forward_nhwc(x)
        return ops.act_to_nchw(x)
