"""Model-level parity for the other BASELINE configs: C1 (PointPillar single), C3 (SECOND + AttFusion, HeterModelBaseline),
C4 (Lift-Splat-Shoot + PointPillars hetero PyramidFusion, bf16)."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import nets, procedural, make_golden, voxelizer, sparse_conv as sc, lss

pytestmark = pytest.mark.gpu

# bf16 mode (bf16 activation storage, fp32 accumulation): every stored tensor carries 2^-9 relative rounding noise and the C2 path
# stores ~60 tensors in sequence, so the MAX abs error over 10^5-10^6 outputs lands at 0.8-2 % of max|ref| (measured, and measured
# the same way for stock PyTorch bf16 autocast of the unmodified reference: bench.py -> cuda_eager_reference).  north_star's
# 1e-2 is met by the fp32-equivalent tc32 mode with three orders of magnitude to spare; the bf16 mode is asserted at 2.5e-2.
BF16_TOL = 2.5e-2

PREC = [("fp32", 1e-3), ("tc32", 1e-3), ("bf16", None)]


def _tol_check(got, ref, tol, name):
    err = (got - ref).abs().max().item()
    scale = max(ref.abs().max().item(), 1.0)
    print(f"{name}: max|ref|={scale:.3f} max_abs_err={err:.3e}")
    assert err < (BF16_TOL if tol is None else tol) * scale, (name, err, scale)


@pytest.fixture(autouse=True)
def _restore_precision():
    from heal_b200 import engine
    old = engine.PRECISION
    yield
    engine.set_precision(old)


def _to_cuda(d):
    if torch.is_tensor(d):
        return d.cuda()
    if isinstance(d, dict):
        return {k: _to_cuda(v) for k, v in d.items()}
    return d


@pytest.mark.parametrize("prec,tol", PREC)
def test_heter_model_baseline_att_vs_reference_golden(golden_dir, prec, tol):
    from heal_b200 import engine
    from heal_b200.models.heter_model_baseline import HeterModelBaseline
    engine.set_precision(prec)
    g = torch.load(os.path.join(golden_dir, "heter_model_baseline_att_small.pt"), weights_only=False)
    m = HeterModelBaseline(copy.deepcopy(g["args"])).eval()
    m.load_state_dict(procedural.make_state_dict(g["shapes"]), strict=True)
    m = m.cuda()
    data = _to_cuda(g["data"])
    data["record_len"] = g["data"]["record_len"]
    with torch.no_grad():
        out = m(data)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _tol_check(out[k].cpu(), g["out"][k], tol, f"{prec}/{k}")


@pytest.mark.parametrize("prec,tol", [("tc32", 1e-3), ("fp32", 1e-3)])
def test_c3_second_attfusion_vs_oracle(prec, tol):
    """BASELINE config 3 (m1m2m3_attfuse.yaml m3 block): SECOND + BaseBEVBackbone(strides 1,2,2, inplanes 128) + AttFusion."""
    from heal_b200 import engine, synth
    from heal_b200.models.heter_model_baseline import HeterModelBaseline
    engine.set_precision(prec)
    rng_ = [-12.8, -12.8, -3, 12.8, 12.8, 1]         # 256 x 256 x 40 voxels @0.1 -> 32 x 32 BEV, 128 channels
    enc_args = {"voxel_size": [0.1, 0.1, 0.1], "lidar_range": rng_, "mean_vfe": {"num_point_features": 4},
                "spconv": {"num_features_in": 4, "num_features_out": 64}, "map2bev": {"feature_num": 128}}
    args = make_golden.baseline_att_args("second", enc_args, inplanes=128, strides=(1, 2, 2))
    args["lidar_range"] = rng_
    m = HeterModelBaseline(copy.deepcopy(args)).eval()
    sd = procedural.make_state_dict(procedural.shapes_of(m))
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    sc_ = synth.scene(23, n_agents=3, rings=32, azimuth=512)
    per_agent = [voxelizer.points_to_voxel_c(p, [0.1, 0.1, 0.1], rng_, 5, 70000) for p in sc_["points"]]
    col = {k: torch.from_numpy(v) for k, v in voxelizer.collate(per_agent).items()}
    pw = torch.from_numpy(sc_["pairwise_t_matrix"]) * 1.0
    pw[..., :2, 3] *= 0.2                               # keep the agents inside the small test range
    dd = {"inputs_m1": col, "agent_modality_list": ["m1"] * 3, "record_len": torch.tensor([3]), "pairwise_t_matrix": pw}
    with torch.no_grad():
        ref = nets.heter_model_baseline(sd, args, dd, encoder_fns={
            "m1": lambda d, mm: sc.second_encoder(sd, "encoder_m1", enc_args, d["inputs_m1"])})
        data = {"inputs_m1": _to_cuda(col), "agent_modality_list": ["m1"] * 3, "record_len": torch.tensor([3]),
                "pairwise_t_matrix": pw.cuda()}
        out = m(data)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _tol_check(out[k].cpu(), ref[k], tol, f"C3/{prec}/{k}")


@pytest.mark.parametrize("prec,tol", [("tc32", 1e-3), ("fp32", 1e-3)])
def test_c1_point_pillar_single_vs_oracle(prec, tol):
    """BASELINE config 1: models/point_pillar.py, 1 agent, 20k-point cloud (range cropped for test time)."""
    from heal_b200 import engine, synth
    from heal_b200.models.point_pillar import PointPillar
    engine.set_precision(prec)
    rng_ = [-51.2, -51.2, -3, 51.2, 51.2, 1]
    args = {"voxel_size": [0.4, 0.4, 4], "lidar_range": rng_, "anchor_number": 2,
            "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
            "point_pillar_scatter": {"num_features": 64},
            "base_bev_backbone": {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
                                  "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]},
            "shrink_header": {"kernal_size": [3], "stride": [1], "padding": [1], "dim": [256], "input_dim": 384},
            "dir_args": {"dir_offset": 0.7853, "num_bins": 2, "anchor_yaw": [0, 90]}}
    m = PointPillar(copy.deepcopy(args)).eval()
    sd = procedural.make_state_dict(procedural.shapes_of(m))
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    pts = synth.lidar_cloud(np.random.default_rng(5), rings=20, azimuth=1000)
    col = {k: torch.from_numpy(v) for k, v in voxelizer.collate([voxelizer.points_to_voxel_c(pts, [0.4, 0.4, 4], rng_, 32, 70000)]).items()}
    with torch.no_grad():
        ref = nets.point_pillar_single(sd, args, {"processed_lidar": col})
        out = m({"processed_lidar": _to_cuda(col)})
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _tol_check(out[k].cpu(), ref[k], tol, f"C1/{prec}/{k}")


@pytest.mark.parametrize("prec,tol", [("bf16", None), ("tc32", 1e-3)])
def test_c4_lss_plus_pointpillar_hetero_vs_oracle(prec, tol):
    """BASELINE config 4 family: agents [m1 (PointPillars), m2 (LSS), m2], camera map zero-padded to the lidar range,
    camera crop mask in PyramidFusion; bf16 tolerance for the bf16 mode."""
    from heal_b200 import engine, synth
    from heal_b200.models.heter_pyramid_collab import HeterPyramidCollab
    engine.set_precision(prec)
    args = make_golden.small_model_args()                  # lidar +-12.8 -> 64x64 pillars, fusion at 32x32
    lcfg = make_golden.lss_small_cfg()
    lcfg["grid_conf"]["xbound"] = [-6.4, 6.4, 0.4]         # camera grid = half the lidar range -> 32x32 @0.4, zero-padded x2
    lcfg["grid_conf"]["ybound"] = [-6.4, 6.4, 0.4]
    lcfg["img_features"] = 64
    args["m2"] = {"core_method": "lift_splat_shoot", "sensor_type": "camera", "encoder_args": lcfg,
                  "camera_mask_args": {"grid_conf": lcfg["grid_conf"]},
                  "backbone_args": {"layer_nums": [3], "layer_strides": [2], "num_filters": [64], "inplanes": 64},
                  "aligner_args": {"core_method": "identity"}}
    m = HeterPyramidCollab(copy.deepcopy(args)).eval()
    sd = procedural.make_state_dict(procedural.shapes_of(m))
    m.load_state_dict(sd, strict=True)
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "heter_pyramid_collab_small.pt"), weights_only=False)
    lidar = {k: v for k, v in g["data"]["inputs_m1"].items()}
    keep = lidar["voxel_coords"][:, 0] == 0                 # agent 0 is the lidar agent
    lidar = {k: v[keep] for k, v in lidar.items()}
    rots, trans, intr, post_rots, post_trans = [torch.from_numpy(a) for a in synth.camera_rig(2, 2, 64, 128)]
    gen = torch.Generator().manual_seed(3)
    imgs = torch.randn(2, 2, 3, 64, 128, generator=gen)
    cam = {"imgs": imgs, "rots": rots, "trans": trans, "intrins": intr, "post_rots": post_rots, "post_trans": post_trans}
    pw = g["data"]["pairwise_t_matrix"]
    aml = ["m1", "m2", "m2"]

    def cam_encoder(dd, mm):                                # oracle LSS on the torch trunk's CPU outputs
        enc = m.encoder_m2
        with torch.no_grad():
            dl, ft = enc.camencode.heads(imgs.view(4, 3, 64, 128))
        fr = lss.create_frustum(lcfg["grid_conf"], lcfg["data_aug_conf"]["final_dim"], lcfg["img_downsample"])
        dx, bx, nx = lss.gen_dx_bx(lcfg["grid_conf"]["xbound"], lcfg["grid_conf"]["ybound"], lcfg["grid_conf"]["zbound"])
        geom = lss.get_geometry(fr, rots, trans, intr, post_rots, post_trans)
        x = lss.outer_product(dl, ft).view(2, 2, -1, *fr.shape[:3]).permute(0, 1, 3, 4, 5, 2)
        return lss.voxel_pooling(geom, x, dx, bx, nx, exact=True)

    def ref_forward():
        # oracle forward with the reference's camera crop/pad + crop mask semantics
        dd = {"inputs_m1": lidar, "agent_modality_list": aml, "record_len": torch.tensor([3]), "pairwise_t_matrix": pw}
        rng_ = args["lidar_range"]
        H, W = rng_[4] - rng_[1], rng_[3] - rng_[0]
        affine = nets.normalize_pairwise_tfm(pw, H, W, 1)
        f1 = nets.resnet_bev_backbone(nets.point_pillar_encoder(sd, "encoder_m1", args["m1"]["encoder_args"], lidar), sd,
                                      "backbone_m1", args["m1"]["backbone_args"])
        f2 = nets.resnet_bev_backbone(cam_encoder(dd, "m2"), sd, "backbone_m2", args["m2"]["backbone_args"])
        import torchvision
        ratio = rng_[3] / lcfg["grid_conf"]["xbound"][1]
        f2 = torchvision.transforms.CenterCrop((int(f2.shape[2] * ratio), int(f2.shape[3] * ratio)))(f2)
        x = torch.stack([f1[0], f2[0], f2[1]])
        info = {"m2": {"crop_ratio_W_m2": ratio, "crop_ratio_H_m2": ratio}}
        fused, occs = nets.pyramid_forward_collab(x, sd, "pyramid_backbone", args["fusion_backbone"], torch.tensor([3]), affine, aml, info)
        fused = nets.downsample_conv(fused, sd, "shrink_conv", args["shrink_header"])
        return {k: nets.conv(fused, sd, k.replace("_preds", "_head")) for k in ("cls_preds", "reg_preds", "dir_preds")}

    with torch.no_grad():
        ref = ref_forward()
        m = m.cuda()
        data = {"inputs_m1": _to_cuda(lidar), "inputs_m2": _to_cuda(cam), "agent_modality_list": aml,
                "record_len": torch.tensor([3]), "pairwise_t_matrix": pw.cuda()}
        out = m(data)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _tol_check(out[k].cpu(), ref[k], tol, f"C4/{prec}/{k}")
