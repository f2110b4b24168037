"""ORACLE — TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference modules (oracle/_ref, see build_ref.py) on the workloads
of workloads/configs.py: on the host cores (`bench.py --impl reference`, bench.py's `cpu_baseline` leg and its full-size parity
check) and, as a baseline leg only, on the GPU through stock PyTorch/cuDNN eager (`cuda_eager_reference`, protocol of
opencood/tools/profiler/params_calc.py:48-79: 50 warm-up + 200 timed iterations between CUDA events).

What is the reference's own code here: every nn.Module on the path (`HeterPyramidCollab`, `PointPillar`, `PillarVFE`,
`PointPillarScatter`, `ResNetBEVBackbone`, `PyramidFusion`, `DownsampleConv`, `warp_affine_simple`, `normalize_pairwise_tfm`).
What is not: voxelisation — spconv is not installable offline, so the restated C voxelizer (oracle/voxelizer.c) produces the
voxel tensors; it is labelled as such wherever a number is reported.  SECOND (config 3) has no spconv-free reference path at all.

Imports neither heal_b200 nor anything that loads libheal_b200.so.
"""
import copy
import os
import time

import numpy as np
import torch

from oracle import ref_shim, voxelizer
from workloads import procedural, configs, synth


def available() -> bool:
    return ref_shim.available()


_MODEL_MODULES = {"heter_pyramid_collab": "HeterPyramidCollab", "point_pillar": "PointPillar",
                  "heter_model_baseline": "HeterModelBaseline"}


def build_model(core_method: str, args: dict, device="cpu"):
    """create_model-style construction (tools/train_utils.py:141-174) of the unmodified class + procedural weights."""
    ref_shim.install()
    import importlib
    lib = importlib.import_module("opencood.models." + core_method)
    cls = getattr(lib, _MODEL_MODULES[core_method])
    model = cls(copy.deepcopy(args)).eval()
    sd = procedural.make_state_dict(procedural.shapes_of(model))
    model.load_state_dict(sd, strict=True)
    return model.to(device), sd


def voxelize_clouds(clouds, voxel_size, lidar_range, max_points, max_voxels):
    """Restated C voxelizer + the reference's collate layout (sp_voxel_preprocessor.py:145-174).  Returns (dict, seconds)."""
    voxelizer.build_c()
    t0 = time.perf_counter()
    per_agent = [voxelizer.points_to_voxel_c(p, voxel_size, lidar_range, max_points, max_voxels) for p in clouds]
    col = {k: torch.from_numpy(v) for k, v in voxelizer.collate(per_agent).items()}
    return col, time.perf_counter() - t0


def c2_data(scene, n_agents, device="cpu"):
    col, tv = voxelize_clouds(scene["clouds"], configs.PILLAR_VOXEL, configs.RANGE, 32, 70000)
    data = {"inputs_m1": {k: v.to(device) for k, v in col.items()}, "agent_modality_list": ["m1"] * n_agents,
            "record_len": torch.tensor([n_agents]), "pairwise_t_matrix": torch.from_numpy(scene["pairwise"]).to(device)}
    return data, tv


def c1_data(cloud, device="cpu"):
    col, tv = voxelize_clouds([cloud], configs.PILLAR_VOXEL, configs.RANGE, 32, 70000)
    return {"processed_lidar": {k: v.to(device) for k, v in col.items()}}, tv


def forward(model, data):
    """The reference forward mutates / consumes its dict: hand it a shallow copy per call."""
    with torch.no_grad():
        return model({k: (dict(v) if isinstance(v, dict) else v) for k, v in data.items()})


def time_cpu_frames(model, datas, warmup, steps):
    for w in range(warmup):
        forward(model, datas[w % len(datas)])
    t0 = time.perf_counter()
    for k in range(steps):
        out = forward(model, datas[k % len(datas)])
    return (time.perf_counter() - t0) / max(steps, 1), out


def time_cuda_eager(model, data, warmup=50, iters=200, allow_tf32=False, autocast_bf16=False):
    """params_calc.py:48-79 protocol.  Returns (ms per frame, outputs of the last frame)."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    torch.backends.cudnn.allow_tf32 = allow_tf32
    torch.backends.cudnn.benchmark = True
    try:
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast_bf16 else torch.autocast("cuda", enabled=False)
        with ctx:
            for _ in range(warmup):
                out = forward(model, data)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                out = forward(model, data)
            b.record()
            torch.cuda.synchronize()
        return a.elapsed_time(b) / iters, out
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = old
