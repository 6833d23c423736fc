"""the north star's target line: one log-likelihood evaluation of the 1 M-cell forest of configs[2] (scaled noise, binomial
division, dt = 15), strict and fast"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gfp_gaussian_process_b200 as ggp
data = ggp.simulate_forest(15873, 6, noise_model="scaled", division_model="binomial", seed=20261018)
f = ggp.Forest(data)
P = ggp.PARAMS_SCALED_BINOMIAL
ll = {}
for mode in ("strict", "fast"):
    f.set_mode(mode)
    ll[mode] = float(ggp.total_likelihood(P, f))
    ms = []
    for _ in range(6):
        ggp.total_likelihood(P, f)
        ms.append(f.last_kernel_ms)
    print("%s: %.3f ms (median of 6), %.3g ctp/s, loglik %.17g, nodes %d, strict reruns %d" % (
        mode, np.median(ms), data.n_ctp / (np.median(ms) * 1e-3), ll[mode], f.last_fast_nodes, f.last_strict_reruns))
print("fast vs strict rel %.2e" % (abs(ll["fast"] - ll["strict"]) / abs(ll["strict"])))
f.close()
