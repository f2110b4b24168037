"""ORACLE — TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by heal_b200/).

CPU fp32 restatement, in plain functional PyTorch over a flat state_dict, of the reference's dense
hot-path modules.  Every function cites the reference code it follows.  Pinned against the
reference's own modules (imported from /root/reference in the build container) by
oracle/make_golden.py -> tests/golden/*.pt and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from collections import Counter
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ---------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------
def bn_eval(x, sd: SD, p: str, eps: float):
    """nn.BatchNorm{1,2}d in eval mode."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    return (x - sd[p + ".running_mean"].view(shape)) / torch.sqrt(sd[p + ".running_var"].view(shape) + eps) \
        * sd[p + ".weight"].view(shape) + sd[p + ".bias"].view(shape)


def conv(x, sd: SD, p: str, stride=1, padding=0, groups=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding, groups=groups)


# ---------------------------------------------------------------------------------------------
# PillarVFE + scatter
# ---------------------------------------------------------------------------------------------
def pillar_vfe(sd: SD, p: str, voxel_features, voxel_num_points, coords, voxel_size, lidar_range):
    """pillar_vfe.py:105-155 (use_norm, use_absolute_xyz, !with_distance, num_filters [C]) + PFNLayer :31-53."""
    vx, vy, vz = voxel_size
    xo, yo, zo = vx / 2 + lidar_range[0], vy / 2 + lidar_range[1], vz / 2 + lidar_range[2]
    vf = voxel_features
    points_mean = vf[:, :, :3].sum(dim=1, keepdim=True) / voxel_num_points.type_as(vf).view(-1, 1, 1)
    f_cluster = vf[:, :, :3] - points_mean
    f_center = torch.zeros_like(vf[:, :, :3])
    f_center[:, :, 0] = vf[:, :, 0] - (coords[:, 3].to(vf.dtype).unsqueeze(1) * vx + xo)
    f_center[:, :, 1] = vf[:, :, 1] - (coords[:, 2].to(vf.dtype).unsqueeze(1) * vy + yo)
    f_center[:, :, 2] = vf[:, :, 2] - (coords[:, 1].to(vf.dtype).unsqueeze(1) * vz + zo)
    feats = torch.cat([vf, f_cluster, f_center], dim=-1)
    T = feats.shape[1]
    mask = (voxel_num_points.int().unsqueeze(1) > torch.arange(T, dtype=torch.int).view(1, -1))
    feats = feats * mask.unsqueeze(-1).type_as(vf)
    x = F.linear(feats, sd[p + ".pfn_layers.0.linear.weight"])
    x = bn_eval(x.permute(0, 2, 1), sd, p + ".pfn_layers.0.norm", 1e-3).permute(0, 2, 1)
    x = F.relu(x)
    x_max = torch.max(x, dim=1, keepdim=True)[0]
    return x_max.squeeze()


def scatter(pillar_features, coords, nx, ny, batch_size=None):
    """point_pillar_scatter.py:19-77."""
    C = pillar_features.shape[1]
    if batch_size is None:
        batch_size = int(coords[:, 0].max().item()) + 1
    out = []
    for b in range(batch_size):
        canvas = torch.zeros(C, nx * ny, dtype=pillar_features.dtype)
        m = coords[:, 0] == b
        tc = coords[m]
        idx = (tc[:, 1] + tc[:, 2] * nx + tc[:, 3]).long()
        canvas[:, idx] = pillar_features[m].t()
        out.append(canvas)
    return torch.stack(out, 0).view(batch_size, C, ny, nx)


# ---------------------------------------------------------------------------------------------
# ResNet / ResNeXt BEV backbones
# ---------------------------------------------------------------------------------------------
def basic_block(x, sd: SD, p: str, stride: int, has_down: bool):
    """resblock.py:18-64, BN eps 1e-5."""
    out = F.relu(bn_eval(conv(x, sd, p + ".conv1", stride, 1), sd, p + ".bn1", 1e-5))
    out = bn_eval(conv(out, sd, p + ".conv2", 1, 1), sd, p + ".bn2", 1e-5)
    idt = x
    if has_down:
        idt = bn_eval(conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1", 1e-5)
    return F.relu(out + idt)


def bottleneck(x, sd: SD, p: str, stride: int, has_down: bool, groups: int):
    """resblock.py:67-122 (expansion forced to 1 by pyramid_fuse.py:72)."""
    out = F.relu(bn_eval(conv(x, sd, p + ".conv1"), sd, p + ".bn1", 1e-5))
    out = F.relu(bn_eval(conv(out, sd, p + ".conv2", stride, 1, groups), sd, p + ".bn2", 1e-5))
    out = bn_eval(conv(out, sd, p + ".conv3"), sd, p + ".bn3", 1e-5)
    idt = x
    if has_down:
        idt = bn_eval(conv(x, sd, p + ".downsample.0", stride, 0), sd, p + ".downsample.1", 1e-5)
    return F.relu(out + idt)


def resnet_modified(x, sd: SD, p: str, kind: str, layer_nums, layer_strides, num_filters, inplanes=64, groups=32):
    """resblock.py:125-219 (_make_layer :178-203, _forward_impl :205-215). Returns the per-level maps."""
    feats = []
    for i, (n, s, planes) in enumerate(zip(layer_nums, layer_strides, num_filters)):
        for b in range(n):
            st = s if b == 0 else 1
            down = (b == 0) and (s != 1 or inplanes != planes)
            bp = f"{p}.layer{i}.{b}"
            x = basic_block(x, sd, bp, st, down) if kind == "basic" else bottleneck(x, sd, bp, st, down, groups)
            if b == 0:
                inplanes = planes
        feats.append(x)
    return feats


def deblocks(xs: List[torch.Tensor], sd: SD, p: str, upsample_strides):
    """base_bev_backbone_resnet.py:54-85 + :127-142: ConvTranspose2d(k=s) + BN(eps 1e-3) + ReLU, concat."""
    ups = []
    for i, x in enumerate(xs):
        s = upsample_strides[i]
        y = F.conv_transpose2d(x, sd[f"{p}.deblocks.{i}.0.weight"], stride=s)
        ups.append(F.relu(bn_eval(y, sd, f"{p}.deblocks.{i}.1", 1e-3)))
    return torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]


def resnet_bev_backbone(x, sd: SD, p: str, cfg):
    """base_bev_backbone_resnet.py:88-109 (ResNetBEVBackbone.forward)."""
    feats = resnet_modified(x, sd, p + ".resnet", "basic", cfg["layer_nums"], cfg["layer_strides"],
                            cfg["num_filters"], cfg.get("inplanes", 64))
    if "upsample_strides" in cfg and len(cfg["upsample_strides"]) > 0:
        return deblocks(feats, sd, p, cfg["upsample_strides"])
    return torch.cat(feats, dim=1) if len(feats) > 1 else feats[0]


def base_bev_backbone(x, sd: SD, p: str, cfg):
    """base_bev_backbone.py:96-124: ZeroPad2d(1)+Conv3x3(s,p=0)+BN(1e-3)+ReLU, k x (Conv3x3+BN+ReLU), deblocks."""
    ups = []
    for i, (n, s) in enumerate(zip(cfg["layer_nums"], cfg["layer_strides"])):
        bp = f"{p}.blocks.{i}"
        x = F.pad(x, (1, 1, 1, 1))
        x = F.relu(bn_eval(conv(x, sd, f"{bp}.1", s, 0), sd, f"{bp}.2", 1e-3))
        for k in range(n):
            x = F.relu(bn_eval(conv(x, sd, f"{bp}.{4 + 3 * k}", 1, 1), sd, f"{bp}.{5 + 3 * k}", 1e-3))
        ups.append(x)
    return deblocks(ups, sd, p, cfg["upsample_strides"])


def downsample_conv(x, sd: SD, p: str, cfg):
    """downsample_conv.py:7-49 (DoubleConv: Conv(k,s,p,bias)+ReLU+Conv3x3(bias)+ReLU)."""
    for i, (k, s, pad) in enumerate(zip(cfg["kernal_size"], cfg["stride"], cfg["padding"])):
        x = F.relu(conv(x, sd, f"{p}.layers.{i}.double_conv.0", s, pad))
        x = F.relu(conv(x, sd, f"{p}.layers.{i}.double_conv.2", 1, 1))
    return x


# ---------------------------------------------------------------------------------------------
# warp + fusion
# ---------------------------------------------------------------------------------------------
def convnext_aligner(x, sd: SD, p: str, num_blocks: int, kernel_size: int = 7):
    """AlignNet(core_method='convnext') = ConvNeXt: `num_blocks` x ConvNeXtBlock
    (feature_alignnet.py:12-39, feature_alignnet_modules.py:299-360; LayerNorm channels_last eps 1e-6 :12-25)."""
    for i in range(num_blocks):
        q = f"{p}.channel_align.model.{i}"
        c = x.shape[1]
        y = F.conv2d(x, sd[q + ".dwconv.weight"], sd[q + ".dwconv.bias"], padding=kernel_size // 2, groups=c)
        y = y.permute(0, 2, 3, 1)
        y = F.layer_norm(y, (c,), sd[q + ".norm.weight"], sd[q + ".norm.bias"], 1e-6)
        y = F.linear(y, sd[q + ".pwconv1.weight"], sd[q + ".pwconv1.bias"])
        y = F.gelu(y)
        y = F.linear(y, sd[q + ".pwconv2.weight"], sd[q + ".pwconv2.bias"])
        if (q + ".gamma") in sd:
            y = sd[q + ".gamma"] * y
        x = x + y.permute(0, 3, 1, 2)
    return x


def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    """utils/transformation_utils.py:68-92."""
    a = pairwise_t_matrix[:, :, :, [0, 1], :][:, :, :, :, [0, 1, 3]].clone()
    a[..., 0, 1] = a[..., 0, 1] * H / W
    a[..., 1, 0] = a[..., 1, 0] * W / H
    a[..., 0, 2] = a[..., 0, 2] / (downsample_rate * discrete_ratio * W) * 2
    a[..., 1, 2] = a[..., 1, 2] / (downsample_rate * discrete_ratio * H) * 2
    return a


def warp_affine_simple(src, M, dsize, align_corners=False):
    """torch_transformation_utils.py:323-332, restated with explicit affine_grid / grid_sample math
    (bilinear, zeros padding): theta stays fp64 through the grid, the grid is cast to src.dtype."""
    B, C, H, W = src.shape
    Ho, Wo = dsize
    th = M.to(torch.float64)

    def lin(n):
        if align_corners:
            return torch.linspace(-1, 1, n, dtype=torch.float64)
        return (torch.arange(n, dtype=torch.float64) * 2 + 1) / n - 1

    xb = lin(Wo).view(1, 1, Wo).expand(B, Ho, Wo)
    yb = lin(Ho).view(1, Ho, 1).expand(B, Ho, Wo)
    gx = (th[:, 0, 0].view(B, 1, 1) * xb + th[:, 0, 1].view(B, 1, 1) * yb + th[:, 0, 2].view(B, 1, 1)).to(src.dtype)
    gy = (th[:, 1, 0].view(B, 1, 1) * xb + th[:, 1, 1].view(B, 1, 1) * yb + th[:, 1, 2].view(B, 1, 1)).to(src.dtype)
    if align_corners:
        ix = (gx + 1) / 2 * (W - 1)
        iy = (gy + 1) / 2 * (H - 1)
    else:
        ix = ((gx + 1) * W - 1) / 2
        iy = ((gy + 1) * H - 1) / 2
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    out = torch.zeros(B, C, Ho, Wo, dtype=src.dtype)
    flat = src.reshape(B, C, H * W)
    for dy, dx in ((0, 0), (0, 1), (1, 0), (1, 1)):
        xi = x0 + dx
        yi = y0 + dy
        wx = (ix - x0) if dx == 1 else (x0 + 1 - ix)
        wy = (iy - y0) if dy == 1 else (y0 + 1 - iy)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long().view(B, 1, Ho * Wo).expand(B, C, Ho * Wo)
        v = torch.gather(flat, 2, idx).view(B, C, Ho, Wo)
        out = out + v * (wx * wy * ok.to(src.dtype)).unsqueeze(1)
    return out


def regroup(x, record_len):
    cum = torch.cumsum(record_len, dim=0)
    return torch.tensor_split(x, cum[:-1].cpu())


def weighted_fuse(x, score, record_len, affine_matrix, align_corners):
    """pyramid_fuse.py:17-63."""
    _, C, H, W = x.shape
    B = affine_matrix.shape[0]
    sx, ss = regroup(x, record_len), regroup(score, record_len)
    out = []
    for b in range(B):
        N = int(record_len[b])
        t = affine_matrix[b][:N, :N]
        f = warp_affine_simple(sx[b], t[0], (H, W), align_corners)
        s = warp_affine_simple(ss[b], t[0], (H, W), align_corners)
        s = s.masked_fill(s == 0, -float("inf"))
        s = torch.softmax(s, dim=0)
        s = torch.where(torch.isnan(s), torch.zeros_like(s), s)
        out.append(torch.sum(f * s, dim=0))
    return torch.stack(out)


def att_fusion(xx, record_len, affine_matrix):
    """fusion_in_one.py:126-151 + ScaledDotProductAttention :41-45."""
    _, C, H, W = xx.shape
    B = affine_matrix.shape[0]
    sx = regroup(xx, record_len)
    out = []
    for b in range(B):
        N = int(record_len[b])
        t = affine_matrix[b][:N, :N]
        x = warp_affine_simple(sx[b], t[0], (H, W))
        n = x.shape[0]
        x = x.view(n, C, -1).permute(2, 0, 1)
        score = torch.bmm(x, x.transpose(1, 2)) / np.sqrt(C)
        attn = F.softmax(score, -1)
        h = torch.bmm(attn, x)
        out.append(h.permute(1, 2, 0).view(n, C, H, W)[0])
    return torch.stack(out)


def pyramid_forward_collab(x, sd: SD, p: str, cfg, record_len, affine, agent_modality_list=None,
                           cam_crop_info=None, training=False):
    """pyramid_fuse.py:104-167 (PyramidFusion.forward_collab), resnext=True."""
    assert cfg["resnext"]
    feats = resnet_modified(x, sd, p + ".resnet", "bottleneck", cfg["layer_nums"], cfg["layer_strides"],
                            cfg["num_filters"], cfg.get("inplanes", 64), groups=32)
    align = cfg.get("align_corners", False)
    fused, occs = [], []
    crop = cam_crop_info is not None and len(cam_crop_info) > 0
    for i, f in enumerate(feats):
        occ = conv(f, sd, f"{p}.single_head_{i}")
        occs.append(occ)
        score = torch.sigmoid(occ) + 1e-4
        if crop and not training:
            mask = torch.ones_like(occ)
            _, _, H, W = mask.shape
            for m in cam_crop_info.keys():
                sel = torch.tensor([1 if a == m else 0 for a in agent_modality_list], dtype=torch.bool)
                crop_H = H / cam_crop_info[m][f"crop_ratio_H_{m}"] - 4
                crop_W = W / cam_crop_info[m][f"crop_ratio_W_{m}"] - 4
                sh, eh = int(H // 2 - crop_H // 2), int(H // 2 + crop_H // 2)
                sw, ew = int(W // 2 - crop_W // 2), int(W // 2 + crop_W // 2)
                mask[sel, :, sh:eh, sw:ew] = 0
                mask[sel] = 1 - mask[sel]
            score = score * mask
        fused.append(weighted_fuse(f, score, record_len, affine, align))
    return deblocks(fused, sd, p, cfg["upsample_strides"]), occs


# ---------------------------------------------------------------------------------------------
# full models
# ---------------------------------------------------------------------------------------------
def point_pillar_encoder(sd: SD, p: str, enc_args, inputs):
    """heter_encoders.py:22-50 (PointPillar encoder)."""
    rng, vs = enc_args["lidar_range"], enc_args["voxel_size"]
    grid = np.round((np.array(rng[3:6]) - np.array(rng[0:3])) / np.array(vs)).astype(np.int64)
    coords = inputs["voxel_coords"]
    pf = pillar_vfe(sd, p + ".pillar_vfe", inputs["voxel_features"], inputs["voxel_num_points"], coords, vs, rng)
    if pf.dim() == 1:
        pf = pf.unsqueeze(0)
    return scatter(pf, coords, int(grid[0]), int(grid[1]))


def heter_pyramid_collab(sd: SD, args, data_dict, encoder_fns=None):
    """heter_pyramid_collab.py:133-209 for lidar (point_pillar) modalities with identity aligners.
    `encoder_fns[modality]` may supply a precomputed encoder feature (used for camera / SECOND)."""
    mods = [k for k in args.keys() if k.startswith("m") and k[1:].isdigit()]
    rng = args["lidar_range"]
    H, W = rng[4] - rng[1], rng[3] - rng[0]
    affine = normalize_pairwise_tfm(data_dict["pairwise_t_matrix"], H, W, 1)
    aml = data_dict["agent_modality_list"]
    cnt = Counter(aml)
    feats = {}
    for m in mods:
        if m not in cnt:
            continue
        if encoder_fns is not None and m in encoder_fns:
            f = encoder_fns[m](data_dict, m)
        else:
            f = point_pillar_encoder(sd, f"encoder_{m}", args[m]["encoder_args"], data_dict[f"inputs_{m}"])
        f = resnet_bev_backbone(f, sd, f"backbone_{m}", args[m]["backbone_args"])
        al = args[m]["aligner_args"]
        if al["core_method"] == "convnext":
            f = convnext_aligner(f, sd, f"aligner_{m}", al["args"]["num_of_blocks"], al["args"].get("kernel_size", 7))
        else:
            assert al["core_method"] == "identity"
        feats[m] = f
    counting = {m: 0 for m in mods}
    lst = []
    for m in aml:
        lst.append(feats[m][counting[m]])
        counting[m] += 1
    x = torch.stack(lst)
    fused, occs = pyramid_forward_collab(x, sd, "pyramid_backbone", args["fusion_backbone"],
                                         data_dict["record_len"], affine, aml, None)
    if "shrink_header" in args:
        fused = downsample_conv(fused, sd, "shrink_conv", args["shrink_header"])
    return {"cls_preds": conv(fused, sd, "cls_head"), "reg_preds": conv(fused, sd, "reg_head"),
            "dir_preds": conv(fused, sd, "dir_head"), "occ_single_list": occs, "fused_feature": fused}


def point_pillar_single(sd: SD, args, data_dict):
    """models/point_pillar.py:52-80 (config C1): VFE -> scatter -> BaseBEVBackbone -> shrink -> heads."""
    rng, vs = args["lidar_range"], args["voxel_size"]
    grid = np.round((np.array(rng[3:6]) - np.array(rng[0:3])) / np.array(vs)).astype(np.int64)
    inp = data_dict["processed_lidar"]
    pf = pillar_vfe(sd, "pillar_vfe", inp["voxel_features"], inp["voxel_num_points"], inp["voxel_coords"], vs, rng)
    if pf.dim() == 1:
        pf = pf.unsqueeze(0)
    x = scatter(pf, inp["voxel_coords"], int(grid[0]), int(grid[1]))
    x = base_bev_backbone(x, sd, "backbone", args["base_bev_backbone"])
    if "shrink_header" in args:
        x = downsample_conv(x, sd, "shrink_conv", args["shrink_header"])
    out = {"cls_preds": conv(x, sd, "cls_head"), "reg_preds": conv(x, sd, "reg_head")}
    if "dir_head.weight" in sd:
        out["dir_preds"] = conv(x, sd, "dir_head")
    return out


def randomize_bn_(sd_or_module, seed=1234):
    """make BN statistics non-trivial so folding bugs show (mean~N(0,.1), var~U(.5,1.5), w~U(.5,1.5), b~N(0,.1))."""
    g = torch.Generator().manual_seed(seed)
    sd = sd_or_module.state_dict() if hasattr(sd_or_module, "state_dict") else sd_or_module
    for k, v in sd.items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            base = k[: -len("running_var")]
            sd[base + "weight"].copy_(torch.rand(v.shape, generator=g) + 0.5)
            sd[base + "bias"].copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def heter_model_baseline(sd: SD, args, data_dict, encoder_fns=None):
    """heter_model_baseline.py:155-236 with fusion_method 'att': encoder -> BaseBEVBackbone -> per-agent shrinker ->
    AttFusion -> (shrink) -> heads.  `encoder_fns[m](data_dict, m)` may supply the encoder output (SECOND / camera)."""
    mods = [k for k in args.keys() if k.startswith("m") and k[1:].isdigit()]
    rng = args["lidar_range"]
    H, W = rng[4] - rng[1], rng[3] - rng[0]
    affine = normalize_pairwise_tfm(data_dict["pairwise_t_matrix"], H, W, 1)
    aml = data_dict["agent_modality_list"]
    cnt = Counter(aml)
    feats = {}
    for m in mods:
        if m not in cnt:
            continue
        if encoder_fns is not None and m in encoder_fns:
            f = encoder_fns[m](data_dict, m)
        else:
            f = point_pillar_encoder(sd, f"encoder_{m}", args[m]["encoder_args"], data_dict[f"inputs_{m}"])
        f = base_bev_backbone(f, sd, f"backbone_{m}", args[m]["backbone_args"])
        feats[m] = downsample_conv(f, sd, f"shrinker_{m}", args[m]["shrink_header"])
    counting = {m: 0 for m in mods}
    lst = []
    for m in aml:
        lst.append(feats[m][counting[m]])
        counting[m] += 1
    x = torch.stack(lst)
    assert args["fusion_method"] == "att"
    fused = att_fusion(x, data_dict["record_len"], affine)
    if "shrink_header" in args:
        fused = downsample_conv(fused, sd, "shrink_conv", args["shrink_header"])
    return {"cls_preds": conv(fused, sd, "cls_head"), "reg_preds": conv(fused, sd, "reg_head"),
            "dir_preds": conv(fused, sd, "dir_head"), "fused_feature": fused}
