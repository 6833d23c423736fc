#!/bin/bash
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r02i.log 2>&1; tail -3 $O/pytest_gpu_r02i.log
python bench.py --impl reference > $O/bench_ref_r02i.json 2> $O/bench_ref_r02i.err
python bench.py > $O/bench_r02i.json 2> $O/bench_r02i.err
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_r02i.json'))
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['pcie']['frac'], d['e2e_resident']['ms_per_step'], d['parity']['fast_gate_met'])
P
