#!/bin/bash
O=gpurun_out
python -m pytest tests/test_correlation.py tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q -k "joint or corr or lag" 2>&1 | tail -3
python tools/r02_jobs/corr_probe.py 3 2>&1 | tee $O/corr_probe_r02c.txt
python - <<'P' 2>&1 | tee $O/walk_probe_r02c.txt
import sys, time, numpy as np
sys.path.insert(0, '.')
import gfp_gaussian_process_b200 as ggp
P2 = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
data = ggp.simulate_forest(1587, 6, params=ggp.PARAMS_SCALED_BINOMIAL, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
f = ggp.Forest(data)
ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
ggp.count_joints(f, P2, 1e-10, 0, 1000)
for _ in range(3):
    n = ggp.count_joints(f, P2, 1e-10); print("count-only walk %.2f ms, %d joints" % (f.last_kernel_ms, n))
t0 = time.perf_counter(); r, c, m, v = ggp.collect_joint_distributions(f, P2, 1e-10, row_begin=0, row_end=100000); dt = time.perf_counter() - t0
print("records: %d in %.3f s = %.3g records/s (walk+sort kernel ms %.2f)" % (len(r), dt, len(r) / dt, f.last_kernel_ms))
P
