"""Eager frames of one BASELINE workload for an ncu launch list (bench.py itself runs warm-up, timed, e2e and instrumented
passes -- several hundred launches more than a launch list needs):

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
      --log-file gpurun_out/launches_c2.csv python profiles/ncu_frame.py --workload c2

The LAST frame is also recorded op by op (ops.PROFILE: family name + number of kernels each C-ABI call launched) into
gpurun_out/opseq_<workload>.json; profiles/derive_traffic.py aligns that sequence with the ncu launch list to get DRAM bytes and
time per kernel family."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--precision", default=None)
    ap.add_argument("--frames", type=int, default=2)
    opt = ap.parse_args()
    from heal_b200 import ops
    prec = opt.precision or os.environ.get("HEAL_PRECISION") or ("bf16" if opt.workload == "c4" else "tc32")
    dev = torch.device("cuda:0")
    wl = bench.GpuWorkload(opt.workload, prec, dev, n_agents=8 if opt.workload == "c5" else None)
    with torch.no_grad():
        for i in range(opt.frames - 1):
            wl.eager(i)
        torch.cuda.synchronize()
        # sentinel the derive script looks for: a 1-element fill right before the recorded frame
        torch.zeros(1, device=dev).fill_(12345.0)
        ops.PROFILE = []
        wl.eager(opt.frames - 1)
        torch.cuda.synchronize()
        recs, ops.PROFILE = ops.PROFILE, None
    seq = [{"family": r[0], "launches": r[5]} for r in recs]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"opseq_{opt.workload}.json"), "w") as fh:
        json.dump({"workload": opt.workload, "precision": prec, "ops": seq}, fh)


if __name__ == "__main__":
    main()
