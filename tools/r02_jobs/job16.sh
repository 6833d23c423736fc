#!/bin/bash
# A/B of the warp-uniform polynomial level: none (per-lane level), always (reduce every step), cond (reduce only if the lanes differ)
for rep in 1 2; do
for v in none always cond; do
  echo "== $v"; GGP_B200_LIB=$PWD/build/ab/libggp_$v.so python tools/fast_probe.py 10000 5 15
  GGP_B200_LIB=$PWD/build/ab/libggp_$v.so python tools/r02_jobs/cfg3_loglik_probe.py | grep "^fast:"
done; done
