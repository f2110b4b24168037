"""normalize_pairwise_tfm mirror (opencood/utils/transformation_utils.py:68-92): fp64 metadata math.
Written with slices instead of list indexing so that it issues no host->device index copies (CUDA-graph capturable)."""
import torch


def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    p = pairwise_t_matrix
    rows = p[..., 0:2, :]                                             # rows [0, 1]
    a = torch.cat([rows[..., 0:2], rows[..., 3:4]], dim=-1).clone()   # cols [0, 1, 3] -> (B,L,L,2,3); input never mutated
    a[..., 0, 1] = a[..., 0, 1] * H / W
    a[..., 1, 0] = a[..., 1, 0] * W / H
    a[..., 0, 2] = a[..., 0, 2] / (downsample_rate * discrete_ratio * W) * 2
    a[..., 1, 2] = a[..., 1, 2] / (downsample_rate * discrete_ratio * H) * 2
    return a
