"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the Lift-Splat-Shoot geometry / outer product / BEV pooling
of opencood/models/heter_encoders.py:83-241 and opencood/utils/camera_utils.py (gen_dx_bx :129-134,
depth_discretization :187-196, QuickCumsum :220-246).  Pinned to the reference's own methods by
oracle/make_golden.py -> tests/golden/lss_small.pt."""
import numpy as np
import torch
import torch.nn.functional as F


def gen_dx_bx(xbound, ybound, zbound):
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.LongTensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def depth_discretization(depth_min, depth_max, num_bins, mode):
    if mode == "UD":
        bin_size = (depth_max - depth_min) / num_bins
        return depth_min + bin_size * np.arange(num_bins)
    if mode == "LID":
        bin_size = 2 * (depth_max - depth_min) / (num_bins * (1 + num_bins))
        return depth_min + bin_size * (np.arange(num_bins) * np.arange(1, 1 + num_bins)) / 2
    raise NotImplementedError


def create_frustum(grid_conf, final_dim, downsample):
    """heter_encoders.py:110-123."""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.tensor(depth_discretization(*grid_conf['ddiscr'], grid_conf['mode']), dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    """heter_encoders.py:125-147 -> (B,N,D,H,W,3)."""
    B, N, _ = trans.shape
    points = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1))
    points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3], points[:, :, :, :, :, 2:3]), 5)
    combine = rots.matmul(torch.inverse(intrins))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
    points = points + trans.view(B, N, 1, 1, 1, 3)
    return points


def cell_index(geom, dx, bx, nx):
    """heter_encoders.py:173-185: truncating index + bounds mask -> (B,N,D,H,W) linear cell or -1."""
    g = ((geom - (bx - dx / 2.)) / dx).long()
    kept = (g[..., 0] >= 0) & (g[..., 0] < nx[0]) & (g[..., 1] >= 0) & (g[..., 1] < nx[1]) & (g[..., 2] >= 0) & (g[..., 2] < nx[2])
    lin = (g[..., 2] * nx[1] + g[..., 1]) * nx[0] + g[..., 0]
    return torch.where(kept, lin, torch.full_like(lin, -1))


def outer_product(depth_logits, feat):
    """lss_submodule.py:228-229 -> (BN, C, D, fH, fW)."""
    depth = F.softmax(depth_logits, dim=1)
    return depth.unsqueeze(1) * feat.unsqueeze(2)


def voxel_pooling(geom, x, dx, bx, nx, exact=False):
    """heter_encoders.py:161-217.  x (B,N,D,H,W,C).  exact=False follows the reference's sort + fp32 cumsum-difference
    (QuickCumsum); exact=True sums each cell in fp64 (what the cumsum trick approximates)."""
    B, N, D, H, W, C = x.shape
    Np = B * N * D * H * W
    x = x.reshape(Np, C)
    g = ((geom - (bx - dx / 2.)) / dx).long().view(Np, 3)
    batch_ix = torch.cat([torch.full([Np // B, 1], ix, dtype=torch.long) for ix in range(B)])
    g = torch.cat((g, batch_ix), 1)
    kept = (g[:, 0] >= 0) & (g[:, 0] < nx[0]) & (g[:, 1] >= 0) & (g[:, 1] < nx[1]) & (g[:, 2] >= 0) & (g[:, 2] < nx[2])
    x, g = x[kept], g[kept]
    final = torch.zeros((B, C, int(nx[2]), int(nx[1]), int(nx[0])), dtype=torch.float64 if exact else x.dtype)
    if exact:
        lin = ((g[:, 3] * nx[2] + g[:, 2]) * nx[1] + g[:, 1]) * nx[0] + g[:, 0]
        flat = torch.zeros((B * int(nx[2]) * int(nx[1]) * int(nx[0]), C), dtype=torch.float64)
        flat.index_add_(0, lin, x.double())
        final = flat.view(B, int(nx[2]), int(nx[1]), int(nx[0]), C).permute(0, 4, 1, 2, 3).float()
    else:
        ranks = g[:, 0] * (nx[1] * nx[2] * B) + g[:, 1] * (nx[2] * B) + g[:, 2] * B + g[:, 3]
        sorts = ranks.argsort()
        x, g, ranks = x[sorts], g[sorts], ranks[sorts]
        x = x.cumsum(0)
        k = torch.ones(x.shape[0], dtype=torch.bool)
        k[:-1] = ranks[1:] != ranks[:-1]
        x, g = x[k], g[k]
        x = torch.cat((x[:1], x[1:] - x[:-1]))
        final[g[:, 3], :, g[:, 2], g[:, 1], g[:, 0]] = x
    return torch.cat(final.unbind(dim=2), 1)
