#!/bin/bash
O=gpurun_out
python -m pytest tests/test_correlation.py tests/test_gpu_fast.py -m gpu -x -q 2>&1 | tail -3
for occ in 3 2; do echo "== 1250 trees, OCC=$occ"; GGP_B200_FAST_OCC=$occ python tools/fast_probe.py 1250 5 15; done 2>&1 | tee $O/fast_small_r02.txt
for occ in 3 2; do echo "== 1250 trees unchunked, OCC=$occ"; GGP_B200_FAST_CHUNKED=0 GGP_B200_FAST_OCC=$occ python tools/fast_probe.py 1250 5 15; done 2>&1 | tee -a $O/fast_small_r02.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_fast5_small_r02.csv python tools/fast_probe.py 1250 5 1 > /dev/null 2>&1
