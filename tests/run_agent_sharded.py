"""torchrun --nproc-per-node N tests/run_agent_sharded.py : N-rank agent-sharded forward (one NCCL all-gather) must equal the
single-GPU forward of the same scene bit for bit.  Run on a multi-GPU box (gpurun --gpus 2)."""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    from heal_b200 import synth, parallel
    from heal_b200.models.heter_pyramid_collab import HeterPyramidCollab
    from oracle import make_golden, procedural
    args = make_golden.small_model_args()
    model = HeterPyramidCollab(copy.deepcopy(args)).eval()
    model.load_state_dict(procedural.make_state_dict(procedural.shapes_of(model)), strict=True)
    model = model.cuda()
    ok = True
    for n_agents in (world, 5, 1):
        sc = synth.scene(5, n_agents=n_agents, max_cav=max(5, n_agents), rings=16, azimuth=256)
        offs = np.concatenate([[0], np.cumsum([p.shape[0] for p in sc["points"]])]).astype(np.int32)
        data = {"inputs_m1": {"points": torch.from_numpy(np.concatenate(sc["points"])).cuda(), "agent_offsets": torch.from_numpy(offs).cuda()},
                "agent_modality_list": ["m1"] * n_agents, "record_len": torch.tensor([n_agents]),
                "pairwise_t_matrix": torch.from_numpy(sc["pairwise_t_matrix"]).cuda()}
        with torch.no_grad():
            ref = model(data)
            out = parallel.forward_agent_sharded(model, data, rank, world)
        torch.cuda.synchronize()
        for k in ("cls_preds", "reg_preds", "dir_preds"):
            same = torch.equal(ref[k], out[k])
            ok = ok and same
            if not same:
                print(f"rank {rank} agents {n_agents} {k}: max diff {(ref[k]-out[k]).abs().max().item():.3e}")
    # graph-captured partition with the row-sharded tail (AgentShardedFrame): same scene, bit-identical heads on every rank
    for n_agents in (world, 5):
        sc = synth.scene(6, n_agents=n_agents, max_cav=max(5, n_agents), rings=16, azimuth=256)
        offs = np.concatenate([[0], np.cumsum([p.shape[0] for p in sc["points"]])]).astype(np.int32)
        pts = torch.from_numpy(np.concatenate(sc["points"])).cuda()
        pw = torch.from_numpy(sc["pairwise_t_matrix"]).cuda()
        data = {"inputs_m1": {"points": pts, "agent_offsets": torch.from_numpy(offs).cuda()},
                "agent_modality_list": ["m1"] * n_agents, "record_len": torch.tensor([n_agents]), "pairwise_t_matrix": pw}
        with torch.no_grad():
            ref = model(data)
            for comm, shard_tail in (("p2p", True), ("p2p", False), ("nccl", True), ("nccl", False)):
                sf = parallel.AgentShardedFrame(model, n_agents, rank, world, 1 << 16, tuple(pw.shape), shard_tail=shard_tail, comm=comm)
                for rep in range(3):                       # p2p mode alternates between two captured graphs / gather buffers
                    sf.load_scene(pts, offs, pw)
                    out = sf.replay()
                    torch.cuda.synchronize()
                    for k in ("cls_preds", "reg_preds", "dir_preds"):
                        same = torch.equal(ref[k], out[k])
                        ok = ok and same
                        if not same:
                            print(f"rank {rank} graph agents {n_agents} comm={comm} shard_tail={shard_tail} rep {rep} {k}: "
                                  f"max diff {(ref[k]-out[k]).abs().max().item():.3e}")
                dist.barrier()
                del sf
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("AGENT_SHARDED_OK" if int(t.item()) == 1 else "AGENT_SHARDED_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
