"""Pre-processor registry entry for the GPU voxelizer.

The reference voxelises inside forked DataLoader workers (opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:62-85),
where CUDA cannot be used.  `GpuVoxelPreprocessor` keeps the reference's pre-processor contract
(`preprocess(pcd_np) -> dict`, `collate_batch(list|dict) -> dict of torch tensors`, registry name via
`preprocess.core_method`) but ships the raw points; the encoders (heal_b200.models.heter_encoders.PointPillar / SECOND)
voxelise on the GPU when `voxel_features` is absent from `inputs_<m>`.
"""
import sys

import numpy as np
import torch


class GpuVoxelPreprocessor:
    def __init__(self, preprocess_params, train):
        self.params = preprocess_params
        self.train = train
        self.lidar_range = self.params['cav_lidar_range']
        self.voxel_size = self.params['args']['voxel_size']
        self.max_points_per_voxel = self.params['args']['max_points_per_voxel']
        self.max_voxels = self.params['args']['max_voxel_train'] if train else self.params['args']['max_voxel_test']
        grid = (np.array(self.lidar_range[3:6]) - np.array(self.lidar_range[0:3])) / np.array(self.voxel_size)
        self.grid_size = np.round(grid).astype(np.int64)

    def preprocess(self, pcd_np):
        return {'points': np.ascontiguousarray(pcd_np[:, :4], dtype=np.float32)}

    def collate_batch(self, batch):
        if isinstance(batch, list):
            clouds = [b['points'] for b in batch]
        elif isinstance(batch, dict):
            clouds = list(batch['points'])
        else:
            sys.exit('Batch has too be a list or a dictionarn')
        offs = np.concatenate([[0], np.cumsum([c.shape[0] for c in clouds])]).astype(np.int32)
        pts = np.concatenate(clouds) if len(clouds) else np.zeros((0, 4), np.float32)
        # the yaml's voxelisation limits travel with the batch: the encoders use them instead of their defaults
        return {'points': torch.from_numpy(pts), 'agent_offsets': torch.from_numpy(offs),
                'agent_offsets_host': offs.tolist(),
                'max_points_per_voxel': int(self.max_points_per_voxel), 'max_voxels': int(self.max_voxels)}
