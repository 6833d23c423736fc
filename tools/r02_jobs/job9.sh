#!/bin/bash
O=gpurun_out
for rep in 1 2; do
for v in pf0 pf1; do
  echo "== $v (table path forced: GGP_B200_FAST_TAB=1)"; GGP_B200_FAST_TAB=1 GGP_B200_LIB=$PWD/build/ab/libggp_$v.so python tools/fast_probe.py 10000 5 15
done; done 2>&1 | tee $O/fast_ab_r02e.txt
echo "== pf0 single-dt path"; GGP_B200_LIB=$PWD/build/ab/libggp_pf0.so python tools/fast_probe.py 10000 5 15 | tee -a $O/fast_ab_r02e.txt
