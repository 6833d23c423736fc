#!/bin/bash
# ncu of the fast kernel's largest launch on the north star's target configuration (1 M cells, scaled / binomial, 6-node rule, degree-15 level)
O=gpurun_out
GGP_B200_FAST_CHUNKED=0 ncu --set full --clock-control none --import-source on -k regex:ggp_fast_loglik --launch-skip 17 --launch-count 1 -f -o $O/prof_fast6_target_r02 python tools/r02_jobs/cfg3_loglik_probe.py > $O/ncu_f_fast6_target_r02.log 2>&1
tail -3 $O/ncu_f_fast6_target_r02.log
