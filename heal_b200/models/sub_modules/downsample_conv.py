"""DownsampleConv / DoubleConv mirror (opencood/models/sub_modules/downsample_conv.py:7-49)."""
import torch.nn as nn

from ... import ops
from ...engine import conv_bn_act


class DoubleConv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1),
            nn.ReLU(inplace=True))

    def forward_nhwc(self, x):
        x = conv_bn_act(x, self.double_conv[0], None, relu=True)
        return conv_bn_act(x, self.double_conv[2], None, relu=True)

    def forward(self, x):
        return ops.act_to_nchw(self.forward_nhwc(ops.to_act(x)))


class DownsampleConv(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layers = nn.ModuleList([])
        input_dim = config['input_dim']
        for (ksize, dim, stride, padding) in zip(config['kernal_size'], config['dim'], config['stride'], config['padding']):
            self.layers.append(DoubleConv(input_dim, dim, kernel_size=ksize, stride=stride, padding=padding))
            input_dim = dim

    def forward_nhwc(self, x):
        for layer in self.layers:
            x = layer.forward_nhwc(x)
        return x

    def forward(self, x):
        return ops.act_to_nchw(self.forward_nhwc(ops.to_act(x)))
