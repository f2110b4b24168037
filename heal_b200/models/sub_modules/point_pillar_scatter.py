"""PointPillarScatter mirror (opencood/models/sub_modules/point_pillar_scatter.py:19-77).

Stand-alone scatter of already computed `pillar_features` (the fused PillarVFE->canvas path used by the
encoders never materialises them).  Output is logically (B, C, ny, nx), physically channels-last."""
import torch
import torch.nn as nn

from ... import ops


class PointPillarScatter(nn.Module):
    def __init__(self, model_cfg):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = model_cfg['num_features']
        self.nx, self.ny, self.nz = [int(v) for v in model_cfg['grid_size']]
        assert self.nz == 1

    def forward(self, batch_dict):
        pf, coords = batch_dict['pillar_features'], batch_dict['voxel_coords']
        if not pf.is_cuda:
            raise RuntimeError("heal_b200 needs CUDA tensors")
        if pf.dim() == 1:
            pf = pf.unsqueeze(0)
        batch_size = int(coords[:, 0].max().item()) + 1      # same host sync as the reference (:45)
        canvas = ops.pillar_scatter(pf.float().contiguous(), coords, self.nx, self.ny, batch_size)      # heal_pillar_scatter
        batch_dict['spatial_features'] = ops.act_to_nchw(canvas)
        return batch_dict
