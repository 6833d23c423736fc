#!/bin/bash
O=gpurun_out
python -m pytest tests/test_correlation.py tests/test_gpu_fast.py -m gpu -x -q 2>&1 | tail -5
python tools/r02_jobs/corr_probe.py 3 2>&1 | tee $O/corr_probe_r02b.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_corr_r02b.csv python tools/r02_jobs/corr_probe.py 1 > $O/ncu_l_corr_r02b.log 2>&1
