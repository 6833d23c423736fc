#!/bin/bash
# One GPU-box round: parity tests, the bench line, the launch list and one full ncu capture of the likelihood kernel's
# largest launch.  usage (from the repo root, through gpurun): bash tools/gpu_round.sh <tag>
# Outputs go to gpurun_out/; copy what should be kept to profiles/ (tools/ncu_summary.py condenses the .ncu-rep).
tag=${1:-round}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -4 gpurun_out/pytest_gpu_$tag.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-300 gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; cut -c1-200 gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ggp_loglik_coop_kernel -s 5 -c 1 -o gpurun_out/prof_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_$tag.log 2>&1
