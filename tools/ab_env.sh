#!/bin/bash
# A/B bench over environment settings: tools/ab_env.sh "VAR=a" "VAR=b OTHER=c" ...
for rep in 1 2; do
for e in "$@"; do
  ms=$(env $e python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.readline()); print("%.3f ms  e2e %.3g" % (d["ms_per_step"], d["e2e"]["value"]))')
  echo "$e  $ms"
done
done
