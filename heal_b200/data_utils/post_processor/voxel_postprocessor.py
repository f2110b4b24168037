"""GPU mirror of opencood/data_utils/post_processor/voxel_postprocessor.py (VoxelPostprocessor, inference side).

Same constructor (`anchor_params`, `train`), `generate_anchor_box()` and `post_process(data_dict, output_dict)` as the reference
(:26-83, :245-405); the decode / filter / rotated-NMS chain runs in one C-ABI call (heal_box_decode_nms, csrc/postprocess.cu) on the
head outputs where they already live -- the reference copies them to the host and loops over shapely polygons.
`generate_label` (training targets) is out of scope and raises."""
from __future__ import annotations

import math

import numpy as np
import torch

from ... import ops


class VoxelPostprocessor:
    def __init__(self, anchor_params, train):
        self.params = anchor_params
        self.train = train
        self.anchor_num = self.params['anchor_args']['num']
        self._buffers = {}
        self._anchors_dev = {}

    # -- voxel_postprocessor.py:30-83 -----------------------------------------------------------
    def generate_anchor_box(self):
        aa = self.params['anchor_args']
        W, H = aa['W'], aa['H']
        r = [math.radians(e) for e in aa['r']]
        assert self.anchor_num == len(r)
        vh, vw = aa['vh'], aa['vw']
        xrange = [aa['cav_lidar_range'][0], aa['cav_lidar_range'][3]]
        yrange = [aa['cav_lidar_range'][1], aa['cav_lidar_range'][4]]
        stride = aa.get('feature_stride', 2)
        x = np.linspace(xrange[0] + vw, xrange[1] - vw, W // stride)
        y = np.linspace(yrange[0] + vh, yrange[1] - vh, H // stride)
        cx, cy = np.meshgrid(x, y)
        cx = np.tile(cx[..., np.newaxis], self.anchor_num)
        cy = np.tile(cy[..., np.newaxis], self.anchor_num)
        cz = np.ones_like(cx) * -1.0
        w, l, h = np.ones_like(cx) * aa['w'], np.ones_like(cx) * aa['l'], np.ones_like(cx) * aa['h']
        r_ = np.ones_like(cx)
        for i in range(self.anchor_num):
            r_[..., i] = r[i]
        if self.params['order'] == 'hwl':
            return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
        if self.params['order'] == 'lhw':
            return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
        raise ValueError('Unknown bbx order.')

    def generate_label(self, **kwargs):
        raise NotImplementedError("heal_b200 covers the inference path; training targets stay in opencood")

    # -- voxel_postprocessor.py:245-405 ---------------------------------------------------------
    def _decode_one(self, cav_content, out):
        cls = out.get('cls_preds', out.get('psm'))
        reg = out.get('reg_preds', out.get('rm'))
        dirp = out.get('dir_preds', out.get('dm'))
        if 'iou_preds' in out:
            raise NotImplementedError("iou_preds rescoring is not part of the GPU post-processor")
        if reg.dim() != 4:
            raise NotImplementedError("anchor-free (CenterPoint) heads are not part of the GPU post-processor")
        dev = cls.device
        anchors = self._device_anchors(cav_content['anchor_box'], dev)
        A, H, W = cls.shape[1], cls.shape[2], cls.shape[3]
        bkey = (H, W, A, str(dev))
        if bkey not in self._buffers:
            self._buffers[bkey] = ops.PostprocessBuffers(H, W, A, 1000, dev)
        dargs = self.params.get('dir_args', {}) if dirp is not None else {}
        return ops.box_decode_nms(cls, reg, dirp, anchors, cav_content['transformation_matrix'],
                                  self.params['target_args']['score_threshold'], self.params['nms_thresh'],
                                  dir_offset=dargs.get('dir_offset', 0.0), num_bins=dargs.get('num_bins', 2),
                                  order=self.params['order'], gt_range=self.params['gt_range'], top=1000, buffers=self._buffers[bkey])

    def _device_anchors(self, anchor, dev):
        """fp32 device copy of the anchor grid, cached per (shape, device).  The reference's collate builds a NEW anchor tensor for
        every batch (intermediate_heter_fusion_dataset.py:704), so neither its data_ptr nor its id identifies the contents: the
        cache keeps the source values and re-uploads when they differ (one small host compare per frame, no unbounded growth)."""
        src = torch.as_tensor(anchor)
        key = (tuple(src.shape), str(dev))
        hit = self._anchors_dev.get(key)
        if hit is not None:
            same = (hit[0] is src) or (hit[0].device == src.device and hit[0].dtype == src.dtype and torch.equal(hit[0], src))
            if same:
                return hit[1]
        devt = src.to(device=dev, dtype=torch.float32).contiguous()
        self._anchors_dev[key] = (src.detach().clone(), devt)
        return devt

    def post_process(self, data_dict, output_dict):
        """Returns (pred_box3d_tensor (K,8,3), scores (K,)) on the device, or (None, None).  One cav (early / intermediate fusion:
        the ego); late fusion's cross-cav NMS stays with the reference."""
        cavs = list(output_dict.keys())
        if len(cavs) != 1:
            raise NotImplementedError("late fusion (NMS across several cavs' boxes) is not part of the GPU post-processor")
        cav = cavs[0]
        assert cav in data_dict
        buf = self._decode_one(data_dict[cav], output_dict[cav])
        # the only host sync of the call: box count + candidate count in ONE device-to-host copy (the reference-shaped return
        # needs the box count; the 4x4 transform is host metadata in the reference's collate and is read before the launch)
        k, above = torch.cat([buf.count, buf.stats[:1]]).tolist()
        if above == 0:                               # nothing above the score threshold (voxel_postprocessor.py:351-352)
            return None, None
        return buf.boxes[:k].clone(), buf.scores[:k].clone()


def make_reference_subclass(ref_cls):
    """Build the class the registry hook installs: a SUBCLASS of the reference's own VoxelPostprocessor (so the datasets keep
    generate_label / generate_gt_bbx / generate_object_center* / collate_batch, which they call on the inference path too:
    intermediate_heter_fusion_dataset.py:179,458,537,666,781; opv2v_basedataset.py:435) whose `post_process` runs on the GPU
    when it can (one cav, anchor-based heads on a CUDA device, no iou_preds) and defers to the reference implementation otherwise
    (late fusion's cross-cav NMS, iou rescoring, CPU tensors)."""
    gpu = VoxelPostprocessor

    class GpuVoxelPostprocessor(ref_cls):
        def __init__(self, anchor_params, train):
            super().__init__(anchor_params, train)
            self._gpu = gpu(anchor_params, train)

        @staticmethod
        def _gpu_eligible(output_dict):
            if len(output_dict) != 1:
                return False
            out = next(iter(output_dict.values()))
            cls = out.get('cls_preds', out.get('psm'))
            reg = out.get('reg_preds', out.get('rm'))
            return ('iou_preds' not in out and torch.is_tensor(cls) and cls.is_cuda and torch.is_tensor(reg) and reg.dim() == 4)

        def post_process(self, data_dict, output_dict):
            if self._gpu_eligible(output_dict):
                return self._gpu.post_process(data_dict, output_dict)
            return super().post_process(data_dict, output_dict)

    GpuVoxelPostprocessor.__name__ = "VoxelPostprocessor"
    return GpuVoxelPostprocessor
