set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ng4.json 2> gpurun_out/bench_ng4.err; cut -c1-330 gpurun_out/bench_ng4.json
GGP_B200_NG4_MIN=1000000000 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ng1.json 2>&1; cut -c1-330 gpurun_out/bench_ng1.json
GGP_B200_NG4_MIN=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ng4all.json 2>&1; cut -c1-330 gpurun_out/bench_ng4all.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ggp_loglik_coop_kernel -s 5 -c 1 -o gpurun_out/prof_r1e -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
