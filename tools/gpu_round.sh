set -x
ncu --set full --clock-control none --import-source on -k regex:ggp_loglik_coop_kernel -s 5 -c 1 -o gpurun_out/prof_r1k_pred -f python tools/measure_configs.py cfg3 > gpurun_out/ncu_cfg3f.log 2>&1
