"""Per-modality encoders with the reference's `(data_dict, modality_name) -> (n,C,H,W)` signature
(opencood/models/heter_encoders.py:22-301).  PointPillar runs PillarVFE + scatter as one kernel."""
import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..engine import require_eval
from .sub_modules.pillar_vfe import PillarVFE
from .sub_modules.point_pillar_scatter import PointPillarScatter


class PointPillar(nn.Module):
    """heter_encoders.py:22-50.  If `inputs_<m>` holds raw points (`points`, `agent_offsets`) instead of
    CPU-voxelised tensors, voxelisation runs on the GPU first (SURVEY.md 7.2 'Boundary vs DataLoader workers')."""

    def __init__(self, args):
        super().__init__()
        grid_size = (np.array(args['lidar_range'][3:6]) - np.array(args['lidar_range'][0:3])) / np.array(args['voxel_size'])
        grid_size = np.round(grid_size).astype(np.int64)
        args['point_pillar_scatter']['grid_size'] = grid_size
        self.pillar_vfe = PillarVFE(args['pillar_vfe'], num_point_features=4, voxel_size=args['voxel_size'],
                                    point_cloud_range=args['lidar_range'])
        self.scatter = PointPillarScatter(args['point_pillar_scatter'])
        self.voxel_size = [float(v) for v in args['voxel_size']]
        self.lidar_range = [float(v) for v in args['lidar_range']]
        self.voxelize_args = args.get('voxelize', {'max_points_per_voxel': 32, 'max_voxels': 70000})

    def forward(self, data_dict, modality_name):
        """reference signature: returns the (n, C, ny, nx) BEV map (logical NCHW, physically channels-last fp32)."""
        return ops.act_to_nchw(self.forward_act(data_dict, modality_name, fmt="f32"))

    def forward_act(self, data_dict, modality_name, fmt=None):
        """internal path: the canvas as an `Act` in the conv engine's activation format (no conversion pass)."""
        from ..engine import act_fmt
        fmt = fmt or act_fmt()
        require_eval(self)
        inp = data_dict[f'inputs_{modality_name}']
        nvox_dev = None
        if 'voxel_features' in inp:
            vf, vc, vn = inp['voxel_features'], inp['voxel_coords'], inp['voxel_num_points']
            batch_size = inp.get('batch_size')
            if batch_size is None:
                batch_size = int(vc[:, 0].max().item()) + 1     # host sync, as the reference (scatter.py:45)
        else:
            pts, offs = inp['points'], inp['agent_offsets']
            vf, vc, vn, nvox_dev = ops.voxelize(pts, offs, self.lidar_range, self.voxel_size,
                                                self.voxelize_args['max_points_per_voxel'],
                                                self.voxelize_args['max_voxels'])
            batch_size = offs.numel() - 1
        w, b = self.pillar_vfe.folded()
        _, canvas = ops.pillar_vfe_scatter(vf, vn, vc, w, b, self.voxel_size, self.lidar_range,
                                           nx=self.scatter.nx, ny=self.scatter.ny, batch_size=batch_size,
                                           num_voxels_dev=nvox_dev, canvas_fmt=fmt)
        return canvas
