"""CUDA-graph capture of one frame.  The per-frame path issues ~80-150 short kernels; replaying them as one graph
removes the Python / launch overhead.  Everything on the raw-points path is capture-safe: no host
syncs (voxel counts stay on the device), caller-independent workspaces, TMA descriptors baked against graph-pool buffers.

The scene shape (number of agents, point-buffer capacity, pairwise matrix shape) is fixed per graph; the live point count
travels in `agent_offsets[-1]` on the device, so clouds of different sizes reuse the same graph.
"""
from __future__ import annotations

import torch

from ._lib import lib


class FrameGraph:
    def __init__(self, model, n_agents: int, point_capacity: int, pairwise_shape, modality: str = "m1", forward_fn=None,
                 device=None, warmup: int = 2):
        dev = device or next(model.parameters()).device
        self.points = torch.zeros((point_capacity, 4), dtype=torch.float32, device=dev)
        self.offsets = torch.zeros((n_agents + 1,), dtype=torch.int32, device=dev)
        self.pairwise = torch.zeros(tuple(pairwise_shape), dtype=torch.float64, device=dev)
        self.capacity = point_capacity
        self.data = {f"inputs_{modality}": {"points": self.points, "agent_offsets": self.offsets},
                     "agent_modality_list": [modality] * n_agents, "record_len": [n_agents],
                     "pairwise_t_matrix": self.pairwise}
        fn = forward_fn or model
        self.pairwise.copy_(torch.eye(4, dtype=torch.float64, device=dev).expand(self.pairwise.shape))
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                fn(self.data)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = lib.heal_launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = fn(self.data)
        self.kernels_per_replay = int(lib.heal_launch_count() - l0)

    def load(self, points: torch.Tensor, offsets: torch.Tensor, pairwise: torch.Tensor):
        """points (P,4) f32, offsets (A+1) i32, pairwise f64 — device or pinned-host tensors (async copies on the current stream)."""
        n = points.shape[0]
        if n > self.capacity:
            raise ValueError(f"scene has {n} points, graph capacity is {self.capacity}")
        self.points[:n].copy_(points, non_blocking=True)
        self.offsets.copy_(offsets, non_blocking=True)
        self.pairwise.copy_(pairwise, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.out
