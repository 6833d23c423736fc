set -x
for v in 3 4; do GGP_B200_COOP_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2>&1; cut -c1-330 gpurun_out/bench_v$v.json; done
GGP_B200_COOP_VARIANT=4 ncu --set full --clock-control none --import-source on -k regex:ggp_loglik_coop_kernel -s 5 -c 1 -o gpurun_out/prof_r1h_v4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f4.log 2>&1
