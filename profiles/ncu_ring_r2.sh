#!/bin/bash
# `ncu --set full` of the row-ring grouped conv (conv3x3_ring.cu): the three 256x256-level launches and the first 128x128-level one
# of one eager C2 frame.  Raw page -> gpurun_out/ncu_r2/c2_ring_raw.csv (copied to profiles/ncu_r2/), summarised by summarise_ncu.py.
set -u
OUT=gpurun_out/ncu_r2
mkdir -p $OUT
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:k_gconv3x3_ring" --launch-count 4 -f -o /tmp/c2_ring \
    python profiles/ncu_frame.py --workload c2 --frames 1 > $OUT/c2_ring.log 2>&1
ncu -i /tmp/c2_ring.ncu-rep --page raw --csv > $OUT/c2_ring_raw.csv 2>> $OUT/c2_ring.log
echo "c2_ring rc=$? rows=$(wc -l < $OUT/c2_ring_raw.csv)"
