#!/bin/bash
for rep in 1 2 3; do
for e in "$@"; do
  ms=$(env $e python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.readline()); print("%.3f ms  e2e %.4g (%.3f ms)" % (d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))')
  echo "$e  $ms"
done
done
