"""AttFusion mirror (opencood/models/fuse_modules/fusion_in_one.py:126-151): same
`forward(x, record_len, affine_matrix)` signature.  The other fusion operators of that file are
out of the hot-path scope."""
import numpy as np
import torch
import torch.nn as nn

from ... import ops


def regroup(x, record_len):
    cum = torch.cumsum(record_len, dim=0)
    return torch.tensor_split(x, cum[:-1].cpu())


def _rl(record_len):
    return [int(v) for v in (record_len.tolist() if torch.is_tensor(record_len) else record_len)]


class ScaledDotProductAttention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.sqrt_dim = np.sqrt(dim)

    def forward(self, query, key, value):
        raise NotImplementedError("executed fused inside AttFusion (heal_att_fuse)")


class AttFusion(nn.Module):
    def __init__(self, feature_dims):
        super().__init__()
        self.att = ScaledDotProductAttention(feature_dims)
        self.feature_dims = feature_dims

    def forward_nhwc(self, x, record_len, affine_matrix):
        """x: Act (sumN,H,W,C) -> Act (B,H,W,C)"""
        rl = _rl(record_len)
        out = ops.act_empty(len(rl), x.H, x.W, x.C, x.fmt, x.device)
        aff = affine_matrix.to(device=x.device, dtype=torch.float64)
        start = 0
        for b, n in enumerate(rl):
            ops.att_fuse(x.images(start, start + n), aff[b, 0, :n].contiguous(), out=out.image(b))
            start += n
        return out

    def forward(self, xx, record_len, affine_matrix):
        return ops.act_to_nchw(self.forward_nhwc(ops.to_act(xx), record_len, affine_matrix))
