set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
python tools/measure_configs.py cfg1 > gpurun_out/configs_r1d.jsonl 2> gpurun_out/configs_r1d.err; cat gpurun_out/configs_r1d.jsonl; tail -3 gpurun_out/configs_r1d.err
