"""Summarise `ncu --set full` raw pages (gpurun_out/ncu_r2/*_raw.csv, produced by profiles/ncu_full_r2.sh) into
profiles/ncu_full_r2_summary.json: per captured launch the duration, DRAM / L2 / tensor-pipe / issue utilisation, achieved
occupancy and the top stall reason -- the counters DESIGN.md cites for the kernels below their roofline."""
import csv
import glob
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_mb", "dram__bytes_write.sum": "dram_write_mb",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_insts",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_subpipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "launch__registers_per_thread": "regs", "launch__grid_size": "grid", "launch__block_size": "block",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
}
UNIT = {"msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}


def load(path):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {"kernel": re.sub(r"\(.*", "", d.get("Kernel Name", ""))[:90]}
        stalls = {}
        for col, unit in zip(hdr, units):
            v = d.get(col, "")
            try:
                x = float(v.replace(",", ""))
            except ValueError:
                continue
            if col in KEEP:
                rec[KEEP[col]] = x * UNIT.get(unit, 1.0)
            elif col.startswith("smsp__average_warp_latency_issue_stalled_") or col.startswith("smsp__average_warps_issue_stalled_"):
                stalls[col.split("stalled_")[1].replace("_per_issue_active.ratio", "").replace(".ratio", "")] = x
        if stalls:
            top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
            rec["top_stalls"] = {k: round(v, 2) for k, v in top}
        out.append(rec)
    return out


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "gpurun_out", "ncu_r2")
    summary = {}
    for f in sorted(glob.glob(os.path.join(src, "*_raw.csv"))):
        summary[os.path.basename(f).replace("_raw.csv", "")] = load(f)
    with open(os.path.join(HERE, "ncu_full_r2_summary.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    for k, v in summary.items():
        print(k, len(v), "launches")
        for r in v[:40]:
            print("   ", {a: (round(b, 2) if isinstance(b, float) else b) for a, b in r.items()})


if __name__ == "__main__":
    main()
