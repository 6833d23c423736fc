#!/bin/bash
# A/B bench of kernel builds: tools/ab_bench.sh build/libggp_a.so build/libggp_b.so ...  (through gpurun; prints ms per step of bench.py for each)
for rep in 1 2; do
for lib in "$@"; do
  ms=$(GGP_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.readline()); print("%.3f ms  e2e %.3g" % (d["ms_per_step"], d["e2e"]["value"]))')
  echo "$lib  $ms"
done
done
