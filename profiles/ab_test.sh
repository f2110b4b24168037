#!/usr/bin/env bash
# Same-box A/B of two library builds (box-to-box variance on this pool is ~5 %, and one toolchain effect cost 30 % on a
# single shape -- see DESIGN.md §9): build the variants here into _ab/lib_<name>.so (git-ignored, shipped by gpurun), then
#   gpurun -- 'bash profiles/ab_test.sh v0 v1 [TC_EXP_ONLY-filter]'
# runs profiles/tc_experiment.py twice per variant, alternating, on the one box the call lands on.
set -euo pipefail
a=${1:?variant A}; b=${2:?variant B}; only=${3:-}
for v in "$a" "$b" "$a" "$b"; do
    cp "_ab/lib_${v}.so" heal_b200/libheal_b200.so
    echo "== ${v}"
    TC_EXP_ONLY="${only}" python profiles/tc_experiment.py 2>&1 | tail -n 12
done
