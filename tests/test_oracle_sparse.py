"""Pins oracle/sparse_conv.py against the dense conv3d formulation of the same operators (the published semantics of
spconv's SubMConv3d / SparseConv3d) on small grids.  CPU only.  (No real-spconv golden exists: parity unpinned there.)"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import sparse_conv as sc


def _random_sparse(rng, batch, shape, n, cin):
    cells = rng.choice(batch * shape[0] * shape[1] * shape[2], size=n, replace=False)
    c = np.zeros((n, 4), dtype=np.int32)
    k = cells.copy()
    c[:, 3] = k % shape[2]; k //= shape[2]
    c[:, 2] = k % shape[1]; k //= shape[1]
    c[:, 1] = k % shape[0]; c[:, 0] = k // shape[0]
    return torch.from_numpy(rng.standard_normal((n, cin)).astype(np.float32)), c


def test_subm_matches_dense_conv3d():
    rng = np.random.default_rng(0)
    shape, B = [9, 12, 14], 2
    feats, coords = _random_sparse(rng, B, shape, 300, 4)
    w = torch.from_numpy(rng.standard_normal((16, 3, 3, 3, 4)).astype(np.float32))
    out = sc.subm_conv3d(feats, coords, w, shape)
    d = sc.dense(feats, coords, shape, B)
    ref = F.conv3d(d, w.permute(0, 4, 1, 2, 3).contiguous(), padding=1)
    c = torch.from_numpy(coords).long()
    torch.testing.assert_close(out, ref[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]], rtol=1e-4, atol=1e-4)


def _check_strided(ksize, stride, pad, shape):
    rng = np.random.default_rng(1)
    B = 2
    feats, coords = _random_sparse(rng, B, shape, 250, 8)
    w = torch.from_numpy(rng.standard_normal((16, *ksize, 8)).astype(np.float32))
    out, oc, oshape = sc.sparse_conv3d(feats, coords, w, shape, stride, pad)
    d = sc.dense(feats, coords, shape, B)
    ref = F.conv3d(d, w.permute(0, 4, 1, 2, 3).contiguous(), stride=stride, padding=pad)
    assert list(ref.shape[2:]) == oshape
    occ = F.conv3d((d.abs().sum(1, keepdim=True) > 0).float() + sc.dense(torch.ones(len(coords), 1), coords, shape, B) * 0,
                   torch.ones(1, 1, *ksize), stride=stride, padding=pad)
    occ = F.conv3d(sc.dense(torch.ones(len(coords), 1), coords, shape, B), torch.ones(1, 1, *ksize), stride=stride, padding=pad) > 0
    got = sc.dense(out, oc, oshape, B)
    assert int(occ.sum()) == len(oc)                       # active set = sites reached by >= 1 active input
    torch.testing.assert_close(got, ref * occ, rtol=1e-4, atol=1e-4)


def test_strided_matches_dense_conv3d():
    _check_strided((3, 3, 3), (2, 2, 2), (1, 1, 1), [9, 12, 14])
    _check_strided((3, 3, 3), (2, 2, 2), (0, 1, 1), [11, 12, 14])     # conv4's asymmetric padding
    _check_strided((3, 1, 1), (2, 1, 1), (0, 0, 0), [5, 12, 14])      # conv_out


def test_height_compression_layout():
    feats = torch.arange(2 * 3, dtype=torch.float32).view(2, 3)
    coords = np.array([[0, 1, 2, 3], [1, 0, 0, 0]], dtype=np.int32)
    bev = sc.height_compression(feats, coords, [2, 4, 5], 2)
    assert bev.shape == (2, 6, 4, 5)
    assert bev[0, 0 * 2 + 1, 2, 3] == 0 and bev[0, 1 * 2 + 1, 2, 3] == 1 and bev[0, 2 * 2 + 1, 2, 3] == 2
    assert bev[1, 0, 0, 0] == 3 and bev[1, 2, 0, 0] == 4
