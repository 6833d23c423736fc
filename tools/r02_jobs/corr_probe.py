"""cfg5 correlation sums: wall and kernel time"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gfp_gaussian_process_b200 as ggp
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
P2 = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
data = ggp.simulate_forest(1587, 6, params=ggp.PARAMS_SCALED_BINOMIAL, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
f = ggp.Forest(data)
ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
ggp.count_joints(f, P2, 1e-10, 0, 1000)
for _ in range(reps):
    t0 = time.perf_counter()
    sums, nj = ggp.api.correlation_sums(f, P2, 15.0, 200)
    print("correlation_sums: %.1f ms wall, %.1f ms device, %d joints, checksum %.17g" % ((time.perf_counter() - t0) * 1e3, f.last_kernel_ms, nj, float(np.asarray(sums, dtype=np.float64)[:, 1:].sum())))
f.close()
