#!/bin/bash
O=gpurun_out
python -m pytest tests/test_gpu_fast.py -x -q 2>&1 | tail -3
for occ in 3 4 2; do GGP_B200_FAST_OCC=$occ python tools/fast_probe.py 10000 5,6 10; done > $O/fast_probe_r02b.txt 2>&1
cat $O/fast_probe_r02b.txt
for c in 1 2 1 2; do GGP_B200_COPY_STREAMS=$c python tools/r02_jobs/e2e_probe.py fast 20; done 2>&1 | tee $O/e2e_copystreams_r02.txt
python tools/r02_jobs/corr_probe.py 3 2>&1 | tee $O/corr_probe_r02a.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_corr_r02a.csv python tools/r02_jobs/corr_probe.py 1 > $O/ncu_l_corr_r02a.log 2>&1
