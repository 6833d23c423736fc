set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for v in 0 2 3; do GGP_B200_COOP_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v$v.json 2>&1; cut -c1-330 gpurun_out/bench_v$v.json; done
for v in 0 2 3; do GGP_B200_COOP_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:ggp_loglik_coop_kernel -s 5 -c 1 -o gpurun_out/prof_r1h_v$v -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f$v.log 2>&1; done
