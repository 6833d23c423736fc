#!/bin/bash
O=gpurun_out
python tools/probes/h2d_probe.py 2>&1 | tee $O/h2d_probe_r02.txt
python -m pytest tests/test_correlation.py -m gpu -x -q 2>&1 | tail -3
python tools/r02_jobs/corr_probe.py 4 2>&1 | tee $O/corr_probe_r02e.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_corr_r02e.csv python tools/r02_jobs/corr_probe.py 1 > $O/ncu_l_corr_r02e.log 2>&1
