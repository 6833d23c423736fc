#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -3
for occ in 3 2; do GGP_B200_FAST_OCC=$occ python tools/fast_probe.py 10000 0,5,6 10; done 2>&1 | tee $O/fast_probe_r02c.txt
GGP_B200_FAST_CHUNKED=0 ncu --set full --clock-control none --import-source on -k regex:ggp_fast_loglik --launch-skip 17 --launch-count 1 -f -o $O/prof_fast5_r02g python tools/fast_probe.py 10000 5 1 > $O/ncu_f_fast5_r02g.log 2>&1
