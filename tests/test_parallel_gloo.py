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
    assert parallel.agent_plan(5, 4) == [[0, 1], [2], [3], [4]] and parallel.agent_plan(7, 3) == [[0, 1, 2], [3, 4], [5, 6]]
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


# ---- graph-captured partition (AgentShardedFrame): symmetric-buffer layout, offset table, tail rows -------------------
def test_tail_rows_partition_covers_the_map_with_halos():
    for world in (2, 4, 8):
        got = []
        for r in range(world):
            t = parallel.tail_rows(256, r, world)
            (r0, r1), (b0, b1), (c0, c1) = t["r"], t["b"], t["c"]
            assert c0 % 4 == 0 and c1 % 4 == 0 and 0 <= c0 <= b0 <= r0 < r1 <= b1 <= c1 <= 256
            assert (b0 == 0 or b0 == r0 - 1) and (b1 == 256 or b1 == r1 + 1)
            assert (c0 == 0 or c0 <= r0 - 2) and (c1 == 256 or c1 >= r1 + 2)          # first 3x3 valid where the second reads it
            got += list(range(r0, r1))
        assert got == list(range(256))
    assert parallel.tail_rows(256, 0, 1) is None and parallel.tail_rows(250, 0, 4) is None


def _worker_layout(rank, world, port, n_agents, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = parallel.agent_plan(n_agents, world)
        slots = len(plan[0])
        shapes = [(8, 8, 16), (4, 4, 32)]
        planes = 2
        foffs, ooffs, chunk = parallel.rank_layout(shapes, planes, slots)
        gen = torch.Generator().manual_seed(99)
        truth_f = [torch.randn(n_agents, planes, h, w, c, generator=gen).to(torch.bfloat16) for (h, w, c) in shapes]
        truth_o = [torch.randn(n_agents, h, w, generator=gen) for (h, w, c) in shapes]
        buf = torch.zeros((world, chunk), dtype=torch.uint8)
        mine = plan[rank]
        for li, (h, w, c) in enumerate(shapes):         # what the conv epilogues do: dense (planes, slots, h, w, c) blocks in MY chunk
            blk = buf[rank][foffs[li]:foffs[li] + planes * slots * h * w * c * 2].view(torch.bfloat16).view(planes, slots, h, w, c)
            occ = buf[rank][ooffs[li]:ooffs[li] + slots * h * w * 4].view(torch.float32).view(slots, h, w)
            for s, a in enumerate(mine):
                blk[:, s] = truth_f[li][a]
                occ[s] = truth_o[li][a]
        dist.all_gather_into_tensor(buf.view(-1), buf[rank].clone())
        table = parallel.agent_offsets_in_gather(plan, shapes, planes, slots)
        flat16, flat32 = buf.view(-1).view(torch.bfloat16), buf.view(-1).view(torch.float32)
        ok = True
        for li, (h, w, c) in enumerate(shapes):
            fo, oo = table[li]
            assert len(fo) == n_agents
            for a in range(n_agents):
                for p in range(planes):
                    got = flat16[fo[a] + p * slots * h * w * c: fo[a] + p * slots * h * w * c + h * w * c].view(h, w, c)
                    ok = ok and torch.equal(got, truth_f[li][a, p])
                ok = ok and torch.equal(flat32[oo[a]: oo[a] + h * w].view(h, w), truth_o[li][a])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_symmetric_buffer_layout_world2():
    for n_agents in (4, 5, 1):
        port = _free_port()
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker_layout, args=(2, port, n_agents, ret), nprocs=2, join=True)
        assert dict(ret) == {0: True, 1: True}, n_agents
