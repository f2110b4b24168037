"""Two eager frames of the C2 workload (HeterPyramidCollab, 5 agents, raw points in) for an ncu launch list:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python profiles/ncu_frame.py
(bench.py itself runs warm-up, timed, e2e and instrumented passes -- several hundred launches more than a launch list needs)."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from heal_b200 import engine
    from heal_b200.models.heter_pyramid_collab import HeterPyramidCollab
    from oracle import procedural
    engine.set_precision(os.environ.get("HEAL_PRECISION", "tc32"))
    dev = torch.device("cuda:0")
    model = HeterPyramidCollab(copy.deepcopy(bench.model_args(max_cav=5))).eval()
    model.load_state_dict(procedural.make_state_dict(procedural.shapes_of(model)), strict=True)
    model = model.to(dev)
    scenes = bench.build_scenes(2, bench.N_AGENTS, seed0=100)
    with torch.no_grad():
        for sc in scenes:
            data = {"inputs_m1": {"points": torch.from_numpy(sc["points"]).to(dev), "agent_offsets": torch.from_numpy(sc["offsets"]).to(dev),
                                  "agent_offsets_host": sc["offsets"].tolist()},
                    "agent_modality_list": ["m1"] * bench.N_AGENTS, "record_len": [bench.N_AGENTS],
                    "pairwise_t_matrix": torch.from_numpy(sc["pairwise"]).to(dev)}
            model(data)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
