#!/bin/bash
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02h.log 2>&1; tail -3 $O/pytest_gpu_r02h.log
python bench.py > $O/bench_r02h.json 2> $O/bench_r02h.err
python bench.py --impl reference > $O/bench_ref_r02h.json 2> $O/bench_ref_r02h.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_r02h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $O/ncu_l_bench_r02h.log 2>&1
GGP_B200_FAST_CHUNKED=0 ncu --set full --clock-control none --import-source on -k regex:ggp_fast_loglik --launch-skip 17 --launch-count 1 -f -o $O/prof_fast5_r02h python tools/fast_probe.py 10000 5 1 > $O/ncu_f_fast5_r02h.log 2>&1
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_r02h.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['pcie'], d['e2e_resident']['ms_per_step'])
print(d['configs']['cfg5_joints'])
P
