#!/bin/bash
O=gpurun_out
python bench.py > $O/bench_r02e.json 2> $O/bench_r02e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_fast5_r02e.csv python tools/fast_probe.py 10000 5 1 > $O/ncu_l_fast5_r02e.log 2>&1
GGP_B200_FAST_CHUNKED=0 ncu --set full --clock-control none --import-source on -k regex:ggp_fast_loglik --launch-skip 17 --launch-count 1 -f -o $O/prof_fast5_r02e python tools/fast_probe.py 10000 5 1 > $O/ncu_f_fast5_r02e.log 2>&1
head -c 300 $O/bench_r02e.json
