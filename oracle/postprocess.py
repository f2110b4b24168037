"""CPU restatement of the reference's detection post-processing:
box decode + score filter + direction-bin fix + corners + projection + size/z filters + rotated NMS + range mask.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke, bench.py's cpu legs) -- never imported by heal_b200/.

Follows, line by line:
  opencood/data_utils/post_processor/voxel_postprocessor.py:30-83   generate_anchor_box
  opencood/data_utils/post_processor/voxel_postprocessor.py:245-405 post_process
  opencood/data_utils/post_processor/voxel_postprocessor.py:408-453 delta_to_boxes3d
  opencood/utils/box_utils.py:152-204  boxes_to_corners_3d,  :278-316 project_box3d,  :693-738 nms_rotated,
  opencood/utils/box_utils.py:840-890  remove_large_pred_bbx / remove_bbx_abnormal_z (incl. the z_len-from-y slip, kept),
  opencood/utils/box_utils.py:384-421  mask_boxes_outside_range_numpy,  opencood/utils/common_utils.py:104-113 limit_period,
  opencood/utils/common_utils.py:230-270 compute_iou / convert_format (shapely Polygon of corners 0..3, x/y).

Pinning: tests/golden/postprocess.pt is produced by the UNMODIFIED reference `VoxelPostprocessor.post_process`
(oracle/make_golden.py).  shapely is not installed in this image, so the reference runs on `QuadPolygon` below as its
`shapely.geometry.Polygon` (convex-quad intersection by Sutherland-Hodgman clipping in float64): the greedy NMS loop, the top-1000
cut, the filters and the decode are pinned to the reference's own code; the polygon-area PRIMITIVE is "parity unpinned" against
GEOS (it is exact convex clipping, GEOS's result differs only by rounding).
"""
from __future__ import annotations

import math

import numpy as np
import torch


# ---- polygon primitive (stand-in for shapely.geometry.Polygon) ---------------------------------------------------------
def _area(pts: np.ndarray) -> float:
    x, y = pts[:, 0], pts[:, 1]
    return 0.5 * float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))


def _clip(subject: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """keep the part of `subject` on the left of the directed edge a->b (counter-clockwise clip polygon)."""
    out = []
    n = len(subject)
    if n == 0:
        return subject
    ex, ey = b[0] - a[0], b[1] - a[1]
    side = lambda p: ex * (p[1] - a[1]) - ey * (p[0] - a[0])
    for i in range(n):
        p, q = subject[i], subject[(i + 1) % n]
        sp, sq = side(p), side(q)
        if sp >= 0:
            out.append(p)
        if (sp > 0 and sq < 0) or (sp < 0 and sq > 0):
            t = sp / (sp - sq)
            out.append(p + t * (q - p))
    return np.array(out, dtype=np.float64).reshape(-1, 2)


def quad_intersection_area(p: np.ndarray, q: np.ndarray) -> float:
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    if (p[:, 0].max() < q[:, 0].min() or q[:, 0].max() < p[:, 0].min() or
            p[:, 1].max() < q[:, 1].min() or q[:, 1].max() < p[:, 1].min()):
        return 0.0                                  # disjoint bounding boxes (exact: the clip would return an empty polygon)
    if _area(p) < 0:
        p = p[::-1]
    if _area(q) < 0:
        q = q[::-1]
    poly = p
    for i in range(4):
        poly = _clip(poly, q[i], q[(i + 1) % 4])
        if len(poly) < 3:
            return 0.0
    return abs(_area(poly))


class _Area:
    def __init__(self, a):
        self.area = a


class QuadPolygon:
    """the three shapely calls the reference makes: Polygon(pts), .intersection(o).area, .union(o).area"""

    def __init__(self, pts):
        self.pts = np.array([(float(x), float(y)) for x, y in pts], dtype=np.float64)
        self.area = abs(_area(self.pts))

    def intersection(self, o):
        return _Area(quad_intersection_area(self.pts, o.pts))

    def union(self, o):
        return _Area(self.area + o.area - quad_intersection_area(self.pts, o.pts))


# ---- restatement -------------------------------------------------------------------------------------------------------
def generate_anchor_box(anchor_args: dict, order: str = "hwl") -> np.ndarray:
    """voxel_postprocessor.py:30-83 -> (H/stride, W/stride, A, 7) float64 [x, y, z, h, w, l, r]"""
    W, H = anchor_args["W"], anchor_args["H"]
    r = [math.radians(e) for e in anchor_args["r"]]
    A = len(r)
    vh, vw = anchor_args["vh"], anchor_args["vw"]
    rng = anchor_args["cav_lidar_range"]
    fs = anchor_args.get("feature_stride", 2)
    x = np.linspace(rng[0] + vw, rng[3] - vw, W // fs)
    y = np.linspace(rng[1] + vh, rng[4] - vh, H // fs)
    cx, cy = np.meshgrid(x, y)
    cx = np.tile(cx[..., np.newaxis], A)
    cy = np.tile(cy[..., np.newaxis], A)
    cz = np.ones_like(cx) * -1.0
    w = np.ones_like(cx) * anchor_args["w"]
    l = np.ones_like(cx) * anchor_args["l"]
    h = np.ones_like(cx) * anchor_args["h"]
    r_ = np.ones_like(cx)
    for i in range(A):
        r_[..., i] = r[i]
    if order == "hwl":
        return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
    if order == "lhw":
        return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
    raise ValueError(order)


def limit_period(val: torch.Tensor, offset=0.5, period=2 * np.pi) -> torch.Tensor:
    return val - torch.floor(val / period + offset) * period          # common_utils.py:112


def delta_to_boxes3d(deltas: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """voxel_postprocessor.py:408-453: (N,14,H,W), (H,W,2,7) -> (N, H*W*2, 7)"""
    N = deltas.shape[0]
    d = deltas.permute(0, 2, 3, 1).contiguous().view(N, -1, 7)
    a = anchors.view(-1, 7).float()
    ad = torch.sqrt(a[:, 4] ** 2 + a[:, 5] ** 2)
    out = torch.zeros_like(d)
    out[..., 0] = d[..., 0] * ad + a[:, 0]
    out[..., 1] = d[..., 1] * ad + a[:, 1]
    out[..., 2] = d[..., 2] * a[:, 3] + a[:, 2]
    out[..., 3:6] = torch.exp(d[..., 3:6]) * a[:, 3:6]
    out[..., 6] = d[..., 6] + a[:, 6]
    return out


def boxes_to_corners_3d(boxes3d: torch.Tensor, order: str) -> torch.Tensor:
    """box_utils.py:152-204"""
    b = boxes3d[:, [0, 1, 2, 5, 4, 3, 6]] if order == "hwl" else boxes3d
    template = b.new_tensor(([1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1],
                             [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1])) / 2
    c = b[:, None, 3:6].repeat(1, 8, 1) * template[None]
    cosa, sina = torch.cos(b[:, 6]), torch.sin(b[:, 6])
    z, o = torch.zeros_like(cosa), torch.ones_like(cosa)
    rot = torch.stack((cosa, sina, z, -sina, cosa, z, z, z, o), dim=1).view(-1, 3, 3).float()
    c = torch.matmul(c.view(-1, 8, 3).float(), rot)
    return c + b[:, None, 0:3]


def project_box3d(c: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """box_utils.py:278-316"""
    h = torch.cat((c.transpose(1, 2), torch.ones((c.shape[0], 1, 8))), dim=1)
    return torch.matmul(T, h)[:, :3, :].transpose(1, 2)


def nms_rotated(boxes: np.ndarray, scores: np.ndarray, threshold: float, top: int = 1000) -> np.ndarray:
    """box_utils.py:693-738 on QuadPolygon"""
    if boxes.shape[0] == 0:
        return np.array([], dtype=np.int32)
    polys = [QuadPolygon([(b[i, 0], b[i, 1]) for i in range(4)]) for b in boxes]
    ixs = scores.argsort()[::-1][:top]
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        iou = np.array([polys[i].intersection(polys[j]).area / polys[i].union(polys[j]).area for j in ixs[1:]], dtype=np.float32)
        remove = np.where(iou > threshold)[0] + 1
        ixs = np.delete(ixs, remove)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


def post_process(cls_preds: torch.Tensor, reg_preds: torch.Tensor, dir_preds, anchors: torch.Tensor, T: torch.Tensor, params: dict):
    """voxel_postprocessor.py:245-405 for one cav (intermediate fusion: the ego).  Returns (boxes (K,8,3) f32, scores (K,) f32)
    or (None, None)."""
    prob = torch.sigmoid(cls_preds.permute(0, 2, 3, 1)).reshape(1, -1)
    batch_box3d = delta_to_boxes3d(reg_preds, anchors)
    mask = torch.gt(prob, params["target_args"]["score_threshold"]).view(1, -1)
    boxes3d = batch_box3d[0][mask[0]]
    scores = prob[0][mask[0]]
    if dir_preds is not None and len(boxes3d) != 0:
        dir_offset, num_bins = params["dir_args"]["dir_offset"], params["dir_args"]["num_bins"]
        dcls = dir_preds.permute(0, 2, 3, 1).contiguous().reshape(1, -1, num_bins)[mask]
        labels = torch.max(dcls, dim=-1)[1]
        period = 2 * np.pi / num_bins
        dir_rot = limit_period(boxes3d[..., 6] - dir_offset, 0, period)
        boxes3d[..., 6] = dir_rot + dir_offset + period * labels.to(dcls.dtype)
        boxes3d[..., 6] = limit_period(boxes3d[..., 6], 0.5, 2 * np.pi)
    if len(boxes3d) == 0:
        return None, None
    corners = project_box3d(boxes_to_corners_3d(boxes3d, params["order"]), T)
    # remove_large_pred_bbx (z_len is taken from the y column and only tested for truthiness, as in the reference) + abnormal z
    x_len = corners[:, :, 0].max(1)[0] - corners[:, :, 0].min(1)[0]
    y_len = corners[:, :, 1].max(1)[0] - corners[:, :, 1].min(1)[0]
    keep = torch.logical_and(torch.logical_and(x_len <= 6, y_len <= 6), y_len)
    keep = torch.logical_and(keep, torch.logical_and(corners[:, :, 2].min(1)[0] >= -3, corners[:, :, 2].max(1)[0] <= 1))
    corners, scores = corners[keep], scores[keep]
    pick = nms_rotated(corners.numpy(), scores.numpy(), params["nms_thresh"])
    corners, scores = corners[pick], scores[pick]
    c = corners.numpy()
    lim = np.array(params["gt_range"], dtype=np.float64)
    m = ((c >= lim[0:3]) & (c <= lim[3:6])).all(axis=2).sum(axis=1) >= 8
    return torch.from_numpy(c[m]), scores[m]
