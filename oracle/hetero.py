"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of HeterPyramidCollab.forward for a mixed lidar + camera scene
(opencood/models/heter_pyramid_collab.py:133-209) on top of oracle/nets.py and oracle/lss.py:

  m1 PointPillars encoder            heter_encoders.py:22-50
  m2 Lift-Splat-Shoot encoder        heter_encoders.py:83-241: CamEncode_Resnet101 trunk + heads (lss_submodule.py:140-233;
                                     the trunk is torchvision's own resnet101 conv1..layer2, as in the reference), depth softmax (x)
                                     features (:227-229), get_geometry (:125-147), voxel_pooling (:161-217, exact segment sum)
  per-modality ResNetBEVBackbone, camera CenterCrop / zero-pad to the lidar range (:153-163), PyramidFusion with the eval-mode
  camera crop mask (pyramid_fuse.py:147-162), shrink header, heads.
"""
import numpy as np
import torch
import torchvision

from oracle import nets, lss


def cam_trunk_heads(sd, p, imgs):
    """CamEncode_Resnet101: imgs (BN,3|4,H,W) -> (depth_logits (BN,D,fH,fW), feat (BN,C,fH,fW)), fp32 CPU."""
    from torchvision.models.resnet import resnet101
    trunk = resnet101(weights=None, zero_init_residual=True).eval()
    own = {}
    for name in ("conv1", "bn1", "layer1", "layer2"):
        for k, v in sd.items():
            if k.startswith(f"{p}.{name}."):
                own[k[len(p) + 1:]] = v
    missing = trunk.load_state_dict(own, strict=False)
    assert not [k for k in missing.unexpected_keys], missing.unexpected_keys
    with torch.no_grad():
        x = imgs[:, :3].clone()
        x = trunk.maxpool(torch.relu(trunk.bn1(trunk.conv1(x))))
        f = trunk.layer2(trunk.layer1(x))
        return nets.conv(f, sd, p + ".depth_head"), nets.conv(f, sd, p + ".image_head")


def lss_encoder(sd, p, enc_args, cam):
    """LiftSplatShoot.forward on CPU.  cam: dict imgs (B,N,3,H,W), rots, trans, intrins, post_rots, post_trans."""
    imgs = cam["imgs"]
    B, N = imgs.shape[:2]
    dl, ft = cam_trunk_heads(sd, p + ".camencode", imgs.reshape(B * N, *imgs.shape[2:]))
    g = enc_args["grid_conf"]
    fr = lss.create_frustum(g, enc_args["data_aug_conf"]["final_dim"], enc_args["img_downsample"])
    dx, bx, nx = lss.gen_dx_bx(g["xbound"], g["ybound"], g["zbound"])
    geom = lss.get_geometry(fr, cam["rots"], cam["trans"], cam["intrins"], cam["post_rots"], cam["post_trans"])
    x = lss.outer_product(dl, ft).view(B, N, -1, *fr.shape[:3]).permute(0, 1, 3, 4, 5, 2)
    return lss.voxel_pooling(geom, x, dx, bx, nx, exact=True)


def heter_pyramid_collab_hetero(sd, args, lidar_inputs, cam_inputs, pairwise, aml):
    """agents in `aml` order, e.g. ['m1','m2','m2']; one scene."""
    rng_ = args["lidar_range"]
    H, W = rng_[4] - rng_[1], rng_[3] - rng_[0]
    n = len(aml)
    with torch.no_grad():
        affine = nets.normalize_pairwise_tfm(pairwise, H, W, 1)
        feats, info = {}, {}
        if "m1" in aml:
            f1 = nets.point_pillar_encoder(sd, "encoder_m1", args["m1"]["encoder_args"], lidar_inputs)
            feats["m1"] = nets.resnet_bev_backbone(f1, sd, "backbone_m1", args["m1"]["backbone_args"])
        if "m2" in aml:
            enc = args["m2"]["encoder_args"]
            f2 = nets.resnet_bev_backbone(lss_encoder(sd, "encoder_m2", enc, cam_inputs), sd, "backbone_m2", args["m2"]["backbone_args"])
            g = args["m2"]["camera_mask_args"]["grid_conf"]
            rw, rh = rng_[3] / g["xbound"][1], rng_[4] / g["ybound"][1]
            f2 = torchvision.transforms.CenterCrop((int(f2.shape[2] * rh), int(f2.shape[3] * rw)))(f2)
            feats["m2"] = f2
            info["m2"] = {"crop_ratio_W_m2": rw, "crop_ratio_H_m2": rh}
        cnt = {m: 0 for m in feats}
        rows = []
        for m in aml:
            rows.append(feats[m][cnt[m]])
            cnt[m] += 1
        x = torch.stack(rows)
        fused, occs = nets.pyramid_forward_collab(x, sd, "pyramid_backbone", args["fusion_backbone"], torch.tensor([n]), affine, aml, info)
        fused = nets.downsample_conv(fused, sd, "shrink_conv", args["shrink_header"])
        return {k: nets.conv(fused, sd, k.replace("_preds", "_head")) for k in ("cls_preds", "reg_preds", "dir_preds")}
