"""configs[4]: records of the first 100 000 start points into pinned host arrays (ggp_joints), repeated"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gfp_gaussian_process_b200 as ggp
P2 = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
data = ggp.simulate_forest(1587, 6, params=ggp.PARAMS_SCALED_BINOMIAL, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
f = ggp.Forest(data)
ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
n = ggp.count_joints(f, P2, 1e-10, 0, 100000)
pin = (torch.empty(n, dtype=torch.int64).pin_memory().numpy(), torch.empty(n, dtype=torch.int64).pin_memory().numpy(),
       torch.empty((n, 44), dtype=torch.float64).pin_memory().numpy())
for _ in range(5):
    t0 = time.perf_counter()
    r, c, m, v = ggp.collect_joint_distributions(f, P2, 1e-10, row_begin=0, row_end=100000, out=pin)
    dt = time.perf_counter() - t0
    print("%d records in %.1f ms = %.3g records/s (walk %.2f ms); checksum %.17g" % (len(r), dt * 1e3, len(r) / dt, f.last_kernel_ms, float(v[::97].sum())))
f.close()
