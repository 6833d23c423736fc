"""configs[4]: the joints walk that only counts (all start points in one launch), device time"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gfp_gaussian_process_b200 as ggp
P2 = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
data = ggp.simulate_forest(1587, 6, params=ggp.PARAMS_SCALED_BINOMIAL, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
f = ggp.Forest(data)
ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
ggp.count_joints(f, P2, 1e-10, 0, 1000)
ms = []
for _ in range(4):
    n = ggp.count_joints(f, P2, 1e-10)
    ms.append(f.last_kernel_ms)
print("lib %s walk blocks/SM %s: count-only walk %.2f ms (min of 4), %d joints" % (os.environ.get("GGP_B200_LIB", "default"), os.environ.get("GGP_B200_WALK_BLOCKS", "1"), min(ms), n))
f.close()
