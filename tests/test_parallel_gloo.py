"""CPU, gloo, world_size 2: the host-side logic of the agent-per-GPU path (agent plan, message packing, the single
all-gather, unpacking in scene order).  The kernels themselves are covered by the -m gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from heal_b200 import parallel


def test_agent_plan_and_layout():
    assert parallel.agent_plan(5, 2) == [[0, 1, 2], [3, 4]]
    assert parallel.agent_plan(8, 8) == [[i] for i in range(8)]
    assert parallel.agent_plan(5, 8) == [[0], [1], [2], [3], [4], [], [], []]
    offs, total = parallel.message_layout([(2, 256, 256, 64), (2, 128, 128, 128), (2, 64, 64, 256)], [(256, 256), (128, 128), (64, 64)])
    assert total == 7426048 * 4 and offs[0] == 0 and offs[3] == (64 * 256 * 256 + 128 * 128 * 128 + 256 * 64 * 64) * 4


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_agents, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = parallel.agent_plan(n_agents, world)
        slots = len(plan[0])
        level_shapes = [(2, 8, 8, 16), (2, 4, 4, 32)]
        occ_shapes = [(8, 8), (4, 4)]
        gen = torch.Generator().manual_seed(1234)
        # every rank can rebuild the whole scene's ground truth; it only PACKS its own agents
        all_lv = [torch.randn(n_agents, *s, generator=gen).to(torch.bfloat16) for s in level_shapes]
        all_oc = [torch.randn(n_agents, *s, generator=gen) for s in occ_shapes]
        mine = plan[rank]
        if mine:
            local = parallel.pack_agents([l[mine] for l in all_lv], [o[mine] for o in all_oc], slots)
        else:
            local = parallel.pack_agents([torch.zeros_like(l[:1]) for l in all_lv], [torch.zeros_like(o[:1]) for o in all_oc], slots)
        gathered = parallel.all_gather_bytes(local, world)
        agent_slots = [r * slots + s for r in range(world) for s in range(len(plan[r]))]
        lv, oc = parallel.unpack_agents(gathered, level_shapes, [torch.bfloat16] * 2, occ_shapes, agent_slots)
        ok = all(torch.equal(a, b) for a, b in zip(lv, all_lv)) and all(torch.equal(a, b) for a, b in zip(oc, all_oc))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(n_agents):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_agents, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_allgather_roundtrip_world2_balanced():
    _run(4)


def test_allgather_roundtrip_world2_ragged_and_idle():
    _run(5)     # 3 + 2 agents
    _run(1)     # rank 1 idle
