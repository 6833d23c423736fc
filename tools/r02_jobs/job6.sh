#!/bin/bash
O=gpurun_out
python tools/r02_jobs/corr_probe.py 5 2>&1 | tee $O/corr_probe_r02d.txt
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02f.log 2>&1; tail -3 $O/pytest_gpu_r02f.log
python bench.py > $O/bench_r02f.json 2> $O/bench_r02f.err
python tools/measure_configs.py cfg1 > $O/configs_cfg1_r02f.jsonl 2>&1
tail -5 $O/configs_cfg1_r02f.jsonl
