#!/bin/bash
# compute-sanitizer pass over the cooperative kernels (named barriers, overlaid scratch regions, cp.async double buffer):
# racecheck + memcheck on the small random-forest parity tests and one joints test, once with the default launch
# selection (1 group per block on small forests) and once with GGP_B200_NG4_MIN=1 (the 4-groups-per-block hot kernels).
# usage (through gpurun, from the repo root): bash tools/sanitizer_round.sh <tag>;  logs -> gpurun_out/sanitizer_*_<tag>.txt
tag=${1:-r02}
SEL='random_forests_all_passes_bitwise or joints_segments_ragged or carry_chain_and_batching'
run() {  # $1 tool, $2 label, rest: env assignments
    local tool=$1 label=$2; shift 2
    env "$@" timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 77 \
        python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitizer_${tool}_${label}_$tag.txt 2>&1
    echo "$tool $label rc=$?"; tail -5 gpurun_out/sanitizer_${tool}_${label}_$tag.txt
}
run memcheck default GGP_DUMMY=0
run memcheck ng4 GGP_B200_NG4_MIN=1
run racecheck default GGP_DUMMY=0
run racecheck ng4 GGP_B200_NG4_MIN=1
