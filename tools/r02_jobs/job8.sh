#!/bin/bash
O=gpurun_out
for rep in 1 2; do
for v in head levels one; do
  echo "== $v"; GGP_B200_LIB=$PWD/build/ab/libggp_$v.so python tools/fast_probe.py 10000 5,6 15
done; done 2>&1 | tee $O/fast_ab_r02d.txt
