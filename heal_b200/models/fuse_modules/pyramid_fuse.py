"""PyramidFusion mirror (opencood/models/fuse_modules/pyramid_fuse.py:17-167).

ResNeXt levels + per-level occupancy heads run on the conv kernels; each level's
warp(feat) / warp(score) / mask / softmax / weighted-sum is ONE kernel per scene
(heal_pyramid_fuse_level) writing straight into the level buffer the deblocks read."""
import torch
import torch.nn as nn

from ... import ops
from ...engine import conv_bn_act, require_eval, act_fmt
from ..sub_modules.base_bev_backbone_resnet import ResNetBEVBackbone
from ..sub_modules.resblock import ResNetModified, Bottleneck


def _record_len_list(record_len):
    if torch.is_tensor(record_len):
        return [int(v) for v in record_len.tolist()]    # host metadata; (B,) int64 on CPU in the reference's collate
    return [int(v) for v in record_len]


def weighted_fuse_nhwc(x, occ, record_len, affine_matrix, align_corners, crop_windows=None):
    """x: Act (sumN,H,W,C); occ (sumN,H,W) f32 logits; affine (B,L,L,2,3) f64 -> Act (B,H,W,C).
    Same contract as weighted_fuse (pyramid_fuse.py:17-63), with score = sigmoid(occ)+1e-4 and the
    crop mask folded in (pyramid_fuse.py:145-162)."""
    rl = _record_len_list(record_len)
    B = len(rl)
    out = ops.act_empty(B, x.H, x.W, x.C, x.fmt, x.device)
    aff = affine_matrix.to(device=x.device, dtype=torch.float64)
    start = 0
    for b, n in enumerate(rl):
        theta = aff[b, 0, :n].contiguous()
        cw = crop_windows[start:start + n] if crop_windows is not None else None
        ops.pyramid_fuse_level(x.images(start, start + n), occ[start:start + n], theta, align_corners, cw, out=out.image(b))
        start += n
    return out


class PyramidFusion(ResNetBEVBackbone):
    def __init__(self, model_cfg, input_channels=64):
        super().__init__(model_cfg, input_channels)
        if model_cfg["resnext"]:
            Bottleneck.expansion = 1
            self.resnet = ResNetModified(Bottleneck, model_cfg['layer_nums'], model_cfg['layer_strides'],
                                         model_cfg['num_filters'], inplanes=model_cfg.get('inplanes', 64),
                                         groups=32, width_per_group=4)
        self.align_corners = model_cfg.get('align_corners', False)
        for i in range(self.num_levels):
            setattr(self, f"single_head_{i}", nn.Conv2d(model_cfg["num_filters"][i], 1, kernel_size=1))

    def _occ_nhwc(self, feat, i, out=None):
        return conv_bn_act(feat, getattr(self, f"single_head_{i}"), None, relu=False, out=out, out_fmt="f32")   # Act f32 (N,H,W,1)

    def forward_single(self, spatial_features):
        require_eval(self)
        feats = self.multiscale_nhwc(ops.to_act(spatial_features))
        occ = [ops.act_to_nchw(self._occ_nhwc(f, i)) for i, f in enumerate(feats)]
        return ops.act_to_nchw(self.decode_nhwc(feats)), occ

    def _crop_windows(self, H, W, agent_modality_list, cam_crop_info, device):
        """(sumN,4) int32 [h0,h1,w0,w1] per agent: region where the score survives (pyramid_fuse.py:147-162)."""
        # static per (map size, modality list): built once and kept on the device (a host->device copy per frame would also be
        # rejected by a CUDA graph capture)
        key = (H, W, tuple(agent_modality_list), str(device))
        cache = self.__dict__.setdefault("_crop_window_cache", {})
        if key in cache:
            return cache[key]

        def clamp(s, e, n):   # python slice semantics of [s:e] on a length-n axis
            s = max(n + s, 0) if s < 0 else min(s, n)
            e = max(n + e, 0) if e < 0 else min(e, n)
            return s, max(e, s)
        win = []
        for m in agent_modality_list:
            if m in cam_crop_info:
                crop_H = H / cam_crop_info[m][f"crop_ratio_H_{m}"] - 4
                crop_W = W / cam_crop_info[m][f"crop_ratio_W_{m}"] - 4
                sh, eh = clamp(int(H // 2 - crop_H // 2), int(H // 2 + crop_H // 2), H)
                sw, ew = clamp(int(W // 2 - crop_W // 2), int(W // 2 + crop_W // 2), W)
                win.append([sh, eh, sw, ew])
            else:
                win.append([0, H, 0, W])
        cache[key] = torch.tensor(win, dtype=torch.int32, device=device)
        return cache[key]

    def forward_collab_nhwc(self, x, record_len, affine_matrix, agent_modality_list=None, cam_crop_info=None):
        """x: Act (sumN,H,W,C) -> (fused Act (B,H,W,sum C_up), [occ Act f32 (sumN,h,w,1)] per level)."""
        require_eval(self)
        feats = self.multiscale_nhwc(x)
        crop = cam_crop_info is not None and len(cam_crop_info) > 0
        fused, occs = [], []
        for i, f in enumerate(feats):
            occ = self._occ_nhwc(f, i)
            occs.append(occ)
            cw = None
            if crop and not self.training:
                cw = self._crop_windows(f.H, f.W, agent_modality_list, cam_crop_info, f.device)
            fused.append(weighted_fuse_nhwc(f, occ.t.view(f.N, f.H, f.W), record_len, affine_matrix, self.align_corners, cw))
        return self.decode_nhwc(fused), occs

    def forward_collab(self, spatial_features, record_len, affine_matrix, agent_modality_list=None, cam_crop_info=None):
        fused, occs = self.forward_collab_nhwc(ops.to_act(spatial_features), record_len, affine_matrix,
                                               agent_modality_list, cam_crop_info)
        return ops.act_to_nchw(fused), [ops.act_to_nchw(o) for o in occs]
