"""Lift-Splat-Shoot kernels vs the oracle and the reference-derived golden."""
import os

import numpy as np
import pytest
import torch

from oracle import lss

pytestmark = pytest.mark.gpu


def _mats(g):
    B, N = g["trans"].shape[:2]
    post_inv = torch.inverse(g["post_rots"]).reshape(B * N, 3, 3)
    combine = g["rots"].matmul(torch.inverse(g["intrins"])).reshape(B * N, 3, 3)
    return B, N, post_inv, combine


def test_lss_cell_index_and_pool_vs_golden(golden_dir):
    from heal_b200 import ops
    g = torch.load(os.path.join(golden_dir, "lss_small.pt"), weights_only=False)
    cfg = g["cfg"]
    fr = lss.create_frustum(cfg["grid_conf"], cfg["data_aug_conf"]["final_dim"], cfg["img_downsample"])
    dx, bx, nx = lss.gen_dx_bx(cfg["grid_conf"]["xbound"], cfg["grid_conf"]["ybound"], cfg["grid_conf"]["zbound"])
    B, N, post_inv, combine = _mats(g)
    ref_cell = lss.cell_index(g["geom"], dx, bx, nx).view(B * N, *fr.shape[:3])
    cell = ops.lss_cell_index(fr.cuda(), post_inv.cuda(), g["post_trans"].reshape(B * N, 3).cuda(), combine.cuda(),
                              g["trans"].reshape(B * N, 3).cuda(), (bx - dx / 2.).tolist(), dx.tolist(), nx.tolist())
    got = cell.cpu().long()
    mism = (got != ref_cell)
    # the reference's own geometry is not reproducible to the last ulp across devices (torch.inverse / matmul), so a
    # frustum point within ~1e-5 cells of a boundary may legitimately land next door; bound the count and check that
    # every mismatch is such a boundary case
    frac = mism.float().mean().item()
    print(f"cell index mismatches: {int(mism.sum())} of {mism.numel()} ({frac:.2e})")
    assert frac < 2e-4
    v = ((g["geom"] - (bx - dx / 2.)) / dx).view(B * N, *fr.shape[:3], 3)
    near = ((v - v.round()).abs() < 1e-3).any(-1)
    assert bool((near | ~mism).all())
    # pooling with the ORACLE's indices (bit-identical inputs): fp tolerance only
    bev = ops.lss_pool(g["depth_logits"].cuda(), g["feat"].cuda(), ref_cell.int().cuda(), N, int(nx[0]), int(nx[1]))
    x = lss.outer_product(g["depth_logits"], g["feat"]).view(B, N, -1, *fr.shape[:3]).permute(0, 1, 3, 4, 5, 2)
    ref_exact = lss.voxel_pooling(g["geom"], x, dx, bx, nx, exact=True)
    got_bev = ops.act_to_nchw(bev).cpu().contiguous()
    torch.testing.assert_close(got_bev, ref_exact, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(got_bev, g["bev"], rtol=1e-3, atol=1e-3)      # vs the unmodified reference (cumsum trick)


def test_lss_full_size_properties():
    """BASELINE config-4 shapes (4 cams, 704x256 -> 32x88, D=48, C=128, 256x256 BEV): mass conservation + linearity."""
    from heal_b200 import ops, synth
    grid_conf = {"xbound": [-51.2, 51.2, 0.4], "ybound": [-51.2, 51.2, 0.4], "zbound": [-10, 10, 20.0], "ddiscr": [2, 50, 48], "mode": "LID"}
    fr = lss.create_frustum(grid_conf, [256, 704], 8)
    dx, bx, nx = lss.gen_dx_bx(grid_conf["xbound"], grid_conf["ybound"], grid_conf["zbound"])
    rots, trans, intr, post_rots, post_trans = [torch.from_numpy(a) for a in synth.camera_rig(2, 4, 256, 704)]
    g = {"rots": rots, "trans": trans, "intrins": intr, "post_rots": post_rots, "post_trans": post_trans}
    B, N, post_inv, combine = _mats(g)
    cell = ops.lss_cell_index(fr.cuda(), post_inv.cuda(), post_trans.reshape(B * N, 3).cuda(), combine.cuda(),
                              trans.reshape(B * N, 3).cuda(), (bx - dx / 2.).tolist(), dx.tolist(), nx.tolist())
    gen = torch.Generator().manual_seed(0)
    logits = torch.randn(B * N, 48, 32, 88, generator=gen).cuda()
    feat = torch.randn(B * N, 128, 32, 88, generator=gen).cuda()
    bev = ops.lss_pool(logits, feat, cell, N, 256, 256).t
    # every in-grid frustum point deposits prob*feat: total mass per agent and channel
    prob = torch.softmax(logits, 1)
    w = (prob * (cell >= 0)).sum(1)                                  # (BN, fH, fW)
    expect = (feat * w.unsqueeze(1)).view(B, N, 128, -1).sum((1, 3))   # (B, C)
    torch.testing.assert_close(bev.sum((1, 2)), expect, rtol=2e-3, atol=2e-2)
    bev2 = ops.lss_pool(logits, 2 * feat, cell, N, 256, 256).t
    torch.testing.assert_close(bev2, 2 * bev, rtol=1e-4, atol=1e-4)   # linear in the features (atomics reorder only)
