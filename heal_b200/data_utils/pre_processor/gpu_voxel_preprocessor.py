"""Pre-processor registry entry for the GPU voxelizer.

The reference voxelises inside forked DataLoader workers (opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:62-85),
where CUDA cannot be used.  `GpuVoxelPreprocessor` keeps the reference's pre-processor contract
(`preprocess(pcd_np) -> dict`, `collate_batch(list|dict) -> dict of torch tensors`, registry name via
`preprocess.core_method`) but ships the raw points; the encoders (heal_b200.models.heter_encoders.PointPillar / SECOND)
voxelise on the GPU when `voxel_features` is absent from `inputs_<m>`.

With `preprocess.args.filter_on_gpu: true` the dataset may also skip its host-side point filters
(shuffle_points / mask_ego_points / mask_points_by_range, pcd_utils.py:41-95): `preprocess` draws the shuffle permutation from
numpy's global RNG exactly as shuffle_points would (same stream position, so a seeded run sees the same order) and the batch
carries `filter_points=True`; the encoders then run heal_mask_points before the voxeliser (ops.raw_points_of).
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
        self.filter_on_gpu = bool(self.params['args'].get('filter_on_gpu', False))
        self.remove_ego = bool(self.params['args'].get('remove_ego', True))

    def preprocess(self, pcd_np):
        out = {'points': np.ascontiguousarray(pcd_np[:, :4], dtype=np.float32)}
        if self.filter_on_gpu:
            out['perm'] = np.random.permutation(pcd_np.shape[0]).astype(np.int32)      # shuffle_points, pcd_utils.py:91-95
        return out

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
        out = {'points': torch.from_numpy(pts), 'agent_offsets': torch.from_numpy(offs),
               'agent_offsets_host': offs.tolist(),
               'max_points_per_voxel': int(self.max_points_per_voxel), 'max_voxels': int(self.max_voxels)}
        if self.filter_on_gpu:
            perms = [b['perm'] for b in batch] if isinstance(batch, list) else list(batch['perm'])
            gperm = np.concatenate([p + o for p, o in zip(perms, offs[:-1])]) if len(perms) else np.zeros((0,), np.int32)
            out.update({'filter_points': True, 'remove_ego': self.remove_ego,
                        'shuffle_perm': torch.from_numpy(gperm.astype(np.int32))})
        return out
