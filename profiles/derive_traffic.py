"""Derive profiles/ncu_traffic_r1.json (DRAM bytes per launch of k_conv2d_tc, share of the frame) from the ncu launch list
profiles/launches_r1_final.csv (last eager frame of `bench.py --no-graph` under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`)."""
import csv
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SCALE = {"msecond": 1e6, "usecond": 1e3, "nsecond": 1.0, "Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def load(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr, data = rows[0], {}
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", "")) * SCALE.get(d["Metric Unit"], 1.0)
        except ValueError:
            continue
        data.setdefault(int(d["ID"]), {"name": d["Kernel Name"]})[d["Metric Name"]] = v
    return [data[i] for i in sorted(data)]


def main():
    src = os.path.join(HERE, "launches_r1_final.csv")
    launches = load(src)
    starts = [i for i, l in enumerate(launches) if "k_cells_insert" in l["name"]]
    frame = launches[starts[-1]:]
    tc = [l for l in frame if "k_conv2d_tc" in l["name"]]
    byt = lambda l: l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    us = lambda l: l.get("gpu__time_duration.sum", 0.0) / 1e3
    out = {
        "source": "profiles/launches_r1_final.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  "--clock-control none, bench.py --no-graph, last eager frame; cold-cache serialised launches)",
        "k_conv2d_tc_launches_per_frame": len(tc),
        "k_conv2d_tc_bytes_per_launch": sum(map(byt, tc)) / max(len(tc), 1),
        "k_conv2d_tc_dram_bytes_per_frame": sum(map(byt, tc)),
        "k_conv2d_tc_us_per_frame_under_ncu": sum(map(us, tc)),
        "k_conv2d_tc_share_of_frame_under_ncu": sum(map(us, tc)) / sum(map(us, frame)),
        "frame_us_under_ncu": sum(map(us, frame)),
        "all_kernels_dram_bytes_per_frame": sum(map(byt, frame)),
    }
    with open(os.path.join(HERE, "ncu_traffic_r1.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
