#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02g.log 2>&1; tail -3 $O/pytest_gpu_r02g.log
SECONDS=0; python bench.py > $O/bench_r02g.json 2> $O/bench_r02g.err; echo "bench wall ${SECONDS}s"
SECONDS=0; python bench.py --impl reference > $O/bench_ref_r02g.json 2> $O/bench_ref_r02g.err; echo "reference arm wall ${SECONDS}s"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_r02g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $O/ncu_l_bench_r02g.log 2>&1
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_r02g.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['e2e_resident']['ms_per_step'])
P
