"""Scene-parallel inference over agents: one process per GPU, rank r owns a contiguous slice of the scene's agents
(voxelize -> encoder -> per-agent backbone -> ResNeXt pyramid levels + occupancy heads), ONE all-gather of the packed
per-agent pyramid (NCCL over NVLink 5 / NVSwitch; gloo in the CPU tests), then every rank runs the fusion, decode,
shrink and heads on the gathered maps (variant B).  The reference has no counterpart: it stacks all
agents on one GPU (intermediate_heter_fusion_dataset.py:619,662) and DDP shards scenes, not agents.

Packed message per agent (bytes): [level0 | level1 | level2 | occ0 | occ1 | occ2] where a level is the agent's feature
map in the conv engine's storage format laid out (planes, H, W, C) and occ is fp32 (h, w).  For the HEAL pyramid at a
256x256 fusion map that is 64*256^2 + 128*128^2 + 256*64^2 (x4 bytes, split-bf16 or fp32) + 86016*4 = 29.7 MB.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def agent_plan(n_agents: int, world: int) -> List[List[int]]:
    """Contiguous, balanced assignment of agents to ranks; every rank gets ceil(n/world) slots (some may be empty)."""
    per = -(-n_agents // world)
    return [[a for a in range(r * per, min((r + 1) * per, n_agents))] for r in range(world)]


def message_layout(level_shapes: Sequence[Tuple[int, int, int, int]], occ_shapes: Sequence[Tuple[int, int]], elem_bytes: int = 2):
    """level_shapes: (planes, H, W, C) per level (bf16 elements; fp32 storage = planes 2 worth of bytes with elem_bytes 2);
    occ_shapes: (h, w) fp32.  Returns (offsets in bytes per segment, total bytes padded to 16)."""
    offs, o = [], 0
    for (p, h, w, c) in level_shapes:
        offs.append(o)
        o += p * h * w * c * elem_bytes
    for (h, w) in occ_shapes:
        offs.append(o)
        o += h * w * 4
    return offs, (o + 15) // 16 * 16


def pack_agents(levels: Sequence[torch.Tensor], occs: Sequence[torch.Tensor], slots: int) -> torch.Tensor:
    """levels[i]: (n_local, planes, H, W, C) any dtype; occs[i]: (n_local, h, w) fp32 -> uint8 (slots, msg_bytes), zero padded."""
    n_local = levels[0].shape[0] if len(levels) else 0
    parts = [l.contiguous().view(torch.uint8).reshape(n_local, -1) for l in levels] + \
            [o.contiguous().view(torch.uint8).reshape(n_local, -1) for o in occs]
    msg = torch.cat(parts, dim=1) if n_local > 0 else None
    nbytes = sum(int(p.shape[1]) for p in parts) if n_local > 0 else None
    if n_local == 0:
        raise ValueError("pack_agents needs at least one local agent (use zero-filled maps for an idle rank)")
    pad = (nbytes + 15) // 16 * 16
    out = torch.zeros((slots, pad), dtype=torch.uint8, device=levels[0].device)
    out[:n_local, :nbytes] = msg
    return out


def unpack_agents(buf: torch.Tensor, level_shapes, level_dtypes, occ_shapes, agent_slots: Sequence[int]):
    """buf: uint8 (total_slots, msg_bytes); agent_slots: slot index of each real agent in scene order.
    Returns (levels [(n, planes, H, W, C)], occs [(n, h, w) fp32])."""
    idx = torch.as_tensor(list(agent_slots), dtype=torch.long, device=buf.device)
    rows = buf.index_select(0, idx)
    levels, occs, o = [], [], 0
    for shp, dt in zip(level_shapes, level_dtypes):
        nb = int(torch.tensor([], dtype=dt).element_size())
        n = 1
        for v in shp:
            n *= v
        levels.append(rows[:, o:o + n * nb].contiguous().view(dt).view(len(agent_slots), *shp))
        o += n * nb
    for (h, w) in occ_shapes:
        occs.append(rows[:, o:o + h * w * 4].contiguous().view(torch.float32).view(len(agent_slots), h, w))
        o += h * w * 4
    return levels, occs


def all_gather_bytes(local: torch.Tensor, world: int) -> torch.Tensor:
    """(slots, B) uint8 per rank -> (world*slots, B): the single collective of the path."""
    out = torch.empty((world * local.shape[0], local.shape[1]), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def forward_agent_sharded(model, data_dict, rank: int, world: int):
    """HeterPyramidCollab.forward with the per-agent part sharded over `world` ranks (single lidar modality 'm1').
    data_dict carries the WHOLE scene on every rank (raw points + agent_offsets); each rank touches only its agents' points."""
    from . import ops
    from .engine import act_fmt
    from .utils.transformation_utils import normalize_pairwise_tfm
    aml = data_dict['agent_modality_list']
    n_agents = len(aml)
    assert all(a == aml[0] for a in aml), "agent-sharded path is implemented for single-modality scenes"
    m = aml[0]
    plan = agent_plan(n_agents, world)
    mine = plan[rank]
    slots = len(plan[0])
    inp = data_dict[f'inputs_{m}']
    offs_host = inp.get('agent_offsets_host')
    if offs_host is None:
        offs_host = inp['agent_offsets'].cpu().tolist()           # one-time host copy when the caller did not provide it
    pb = model.pyramid_backbone
    enc, bb = getattr(model, f"encoder_{m}"), getattr(model, f"backbone_{m}")
    fmt = act_fmt()
    # ---- local agents: encoder -> backbone -> ResNeXt levels + occ heads --------------------------------
    idle = len(mine) == 0
    use = mine if not idle else [0]                                # an idle rank computes agent 0 and discards it
    lo, hi = offs_host[use[0]], offs_host[use[-1] + 1]
    sub_offs = torch.tensor([o - lo for o in offs_host[use[0]:use[-1] + 2]], dtype=torch.int32, device=inp['points'].device)
    sub = {f'inputs_{m}': {'points': inp['points'][lo:hi], 'agent_offsets': sub_offs}}
    x = enc.forward_act(sub, m)
    x = bb.decode_nhwc(bb.multiscale_nhwc(x))
    feats = pb.multiscale_nhwc(x)
    occs = [pb._occ_nhwc(f, i) for i, f in enumerate(feats)]
    def per_agent(a: "ops.Act"):
        t = a.t if a.fmt != "f32" else a.t.unsqueeze(0)
        return t.permute(1, 0, 2, 3, 4).contiguous()               # (n_local, planes|1, H, W, C)
    lv = [per_agent(f) for f in feats]
    oc = [o.t.view(o.N, o.H, o.W) for o in occs]
    level_shapes = [tuple(l.shape[1:]) for l in lv]
    level_dtypes = [l.dtype for l in lv]
    occ_shapes = [tuple(o.shape[1:]) for o in oc]
    if idle:
        lv = [torch.zeros_like(l) for l in lv]
        oc = [torch.zeros_like(o) for o in oc]
    local = pack_agents(lv, oc, slots)
    # ---- the one exchange step -------------------------------------------------------------------------------
    with ops._Prof("allgather_bev_pyramid"):
        gathered = all_gather_bytes(local, world) if world > 1 else local
    agent_slots = [r * slots + s for r in range(world) for s in range(len(plan[r]))]
    g_levels, g_occs = unpack_agents(gathered, level_shapes, level_dtypes, occ_shapes, agent_slots)
    # ---- replicated tail: fuse x3 -> decode -> shrink -> heads ------------------------------------------------
    rng = model.cav_range
    affine = normalize_pairwise_tfm(data_dict['pairwise_t_matrix'], model.H, model.W, model.fake_voxel_size)
    from .models.fuse_modules.pyramid_fuse import weighted_fuse_nhwc
    fused = []
    for lvl, occ in zip(g_levels, g_occs):
        t = lvl.permute(1, 0, 2, 3, 4).contiguous()                 # (planes, n, H, W, C)
        act = ops.Act(t if fmt != "f32" else t[0], fmt)
        fused.append(weighted_fuse_nhwc(act, occ.contiguous(), data_dict['record_len'], affine, pb.align_corners, None))
    f = pb.decode_nhwc(fused)
    if model.shrink_flag:
        f = model.shrink_conv.forward_nhwc(f)
    cls, reg, dirp = model._heads(f)
    return {'pyramid': 'collab', 'cls_preds': cls, 'reg_preds': reg, 'dir_preds': dirp,
            'occ_single_list': [o.unsqueeze(1) for o in g_occs]}
