"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the spconv operators the reference's SECOND encoder uses
(opencood/models/sub_modules/sparse_backbone_3d.py:11-152, height_compression.py:10-27).

spconv is a third-party dependency that is NOT vendored, NOT pinned (README.md:107-116) and NOT installed here, so this
restates its published semantics: SubMConv3d / SparseConv3d are torch.nn.functional.conv3d (cross-correlation, no
kernel flip) on the densified input, evaluated only at ACTIVE output sites; SubM keeps the input's active set, a
regular sparse conv activates every output reached by an active input, out = floor((in + 2p - k)/s) + 1; BatchNorm1d
and ReLU act on the active rows; `.dense()` is zero elsewhere.  Weight layout = spconv 2.x (Cout, kz, ky, kx, Cin).
PARITY UNPINNED against real spconv (no build, no golden vectors exist); it IS pinned against the dense conv3d
formulation by tests/test_oracle_sparse.py on cropped grids.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _lin(c, shape):
    return ((c[:, 0].astype(np.int64) * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def _lookup(sorted_keys, order, q, valid):
    idx = np.searchsorted(sorted_keys, q)
    idx = np.minimum(idx, len(sorted_keys) - 1)
    hit = valid & (sorted_keys[idx] == q)
    return np.where(hit, order[idx], -1)


def subm_conv3d(feats, coords, weight, spatial_shape):
    """feats (M,Cin) f32 tensor, coords (M,4) int [b,z,y,x] numpy, weight (Cout,kz,ky,kx,Cin)."""
    coords = np.asarray(coords)
    keys = _lin(coords, spatial_shape)
    order = np.argsort(keys)
    sk = keys[order]
    cout, kz, ky, kx, cin = weight.shape
    out = torch.zeros((feats.shape[0], cout), dtype=feats.dtype)
    for dz in range(kz):
        for dy in range(ky):
            for dx in range(kx):
                n = coords.copy()
                n[:, 1] += dz - kz // 2
                n[:, 2] += dy - ky // 2
                n[:, 3] += dx - kx // 2
                ok = (n[:, 1] >= 0) & (n[:, 1] < spatial_shape[0]) & (n[:, 2] >= 0) & (n[:, 2] < spatial_shape[1]) & \
                     (n[:, 3] >= 0) & (n[:, 3] < spatial_shape[2])
                src = _lookup(sk, order, _lin(np.where(ok[:, None], n, 0), spatial_shape), ok)
                m = src >= 0
                if m.any():
                    rows = torch.from_numpy(np.nonzero(m)[0])
                    out[rows] += feats[torch.from_numpy(src[m])] @ weight[:, dz, dy, dx, :].t()
    return out


def out_shape(spatial_shape, ksize, stride, padding):
    return [(spatial_shape[i] + 2 * padding[i] - ksize[i]) // stride[i] + 1 for i in range(3)]


def sparse_conv3d(feats, coords, weight, spatial_shape, stride, padding):
    """Returns (feats_out (Mo,Cout), coords_out (Mo,4) sorted by linear site index, out_spatial_shape)."""
    coords = np.asarray(coords)
    cout, kz, ky, kx, cin = weight.shape
    oshape = out_shape(spatial_shape, (kz, ky, kx), stride, padding)
    cand = []
    for dz in range(kz):
        for dy in range(ky):
            for dx in range(kx):
                nz, ny, nx = coords[:, 1] + padding[0] - dz, coords[:, 2] + padding[1] - dy, coords[:, 3] + padding[2] - dx
                ok = (nz >= 0) & (ny >= 0) & (nx >= 0) & (nz % stride[0] == 0) & (ny % stride[1] == 0) & (nx % stride[2] == 0)
                oz, oy, ox = nz // stride[0], ny // stride[1], nx // stride[2]
                ok &= (oz < oshape[0]) & (oy < oshape[1]) & (ox < oshape[2])
                o = np.stack([coords[:, 0], oz, oy, ox], 1)
                cand.append((ok, o, (dz, dy, dx)))
    allk = np.concatenate([_lin(o[ok], oshape) for ok, o, _ in cand])
    uk = np.unique(allk)
    out = torch.zeros((len(uk), cout), dtype=feats.dtype)
    for ok, o, (dz, dy, dx) in cand:
        if not ok.any():
            continue
        rows = np.searchsorted(uk, _lin(o[ok], oshape))
        src = np.nonzero(ok)[0]
        out.index_add_(0, torch.from_numpy(rows), feats[torch.from_numpy(src)] @ weight[:, dz, dy, dx, :].t())
    oc = np.zeros((len(uk), 4), dtype=np.int32)
    k = uk.copy()
    oc[:, 3] = k % oshape[2]; k //= oshape[2]
    oc[:, 2] = k % oshape[1]; k //= oshape[1]
    oc[:, 1] = k % oshape[0]; oc[:, 0] = k // oshape[0]
    return out, oc, oshape


def bn_relu(x, sd, p, eps=1e-3):
    x = (x - sd[p + ".running_mean"]) / torch.sqrt(sd[p + ".running_var"] + eps) * sd[p + ".weight"] + sd[p + ".bias"]
    return F.relu(x)


def dense(feats, coords, spatial_shape, batch):
    C = feats.shape[1]
    out = torch.zeros((batch, C, *spatial_shape), dtype=feats.dtype)
    c = torch.from_numpy(np.asarray(coords)).long()
    out[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feats
    return out


def voxel_backbone8x(sd, p, voxel_features, coords, batch_size, grid_size_xyz):
    """sparse_backbone_3d.py:33-152.  Returns (feats, coords, spatial_shape) of `encoded_spconv_tensor`."""
    shape = [int(grid_size_xyz[2]) + 1, int(grid_size_xyz[1]), int(grid_size_xyz[0])]     # grid_size[::-1] + [1,0,0]
    c = np.asarray(coords)
    x = bn_relu(subm_conv3d(voxel_features, c, sd[p + ".conv_input.0.weight"], shape), sd, p + ".conv_input.1")
    x = bn_relu(subm_conv3d(x, c, sd[p + ".conv1.0.0.weight"], shape), sd, p + ".conv1.0.1")
    for name, stride, pad in (("conv2", (2, 2, 2), (1, 1, 1)), ("conv3", (2, 2, 2), (1, 1, 1)), ("conv4", (2, 2, 2), (0, 1, 1))):
        x, c, shape = sparse_conv3d(x, c, sd[f"{p}.{name}.0.0.weight"], shape, stride, pad)
        x = bn_relu(x, sd, f"{p}.{name}.0.1")
        for j in (1, 2):
            x = bn_relu(subm_conv3d(x, c, sd[f"{p}.{name}.{j}.0.weight"], shape), sd, f"{p}.{name}.{j}.1")
    x, c, shape = sparse_conv3d(x, c, sd[p + ".conv_out.0.weight"], shape, (2, 1, 1), (0, 0, 0))
    x = bn_relu(x, sd, p + ".conv_out.1")
    return x, c, shape


def height_compression(feats, coords, spatial_shape, batch):
    """height_compression.py:21-23: .dense() -> (N,C,D,H,W) -> view (N, C*D, H, W)."""
    d = dense(feats, coords, spatial_shape, batch)
    N, C, D, H, W = d.shape
    return d.view(N, C * D, H, W)


def mean_vfe(voxel_features, voxel_num_points):
    """mean_vfe.py:26-30."""
    s = voxel_features.sum(dim=1)
    return (s / torch.clamp_min(voxel_num_points.view(-1, 1), min=1.0).type_as(voxel_features)).contiguous()


def second_encoder(sd, p, enc_args, inputs):
    """heter_encoders.py:52-81 (SECOND): MeanVFE -> VoxelBackBone8x -> HeightCompression."""
    rng, vs = np.array(enc_args["lidar_range"]), np.array(enc_args["voxel_size"])
    grid = np.round((rng[3:6] - rng[:3]) / vs).astype(np.int64)
    coords = inputs["voxel_coords"].numpy() if torch.is_tensor(inputs["voxel_coords"]) else inputs["voxel_coords"]
    batch = int(coords[:, 0].max()) + 1
    vf = mean_vfe(inputs["voxel_features"], inputs["voxel_num_points"])
    x, c, shape = voxel_backbone8x(sd, p + ".spconv_block", vf, coords, batch, grid)
    return height_compression(x, c, shape, batch)
