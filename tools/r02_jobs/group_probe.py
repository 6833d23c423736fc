"""per-call overhead of the group handle: one fast evaluation over k shards on the visible devices (wrapping around)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gfp_gaussian_process_b200 as ggp
nd = torch.cuda.device_count()
trees = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
P = ggp.PARAMS_CONST_GAUSS
d = ggp.simulate_forest(trees, 6, seed=20261018)
for shards in (1, 2, 4, 8):
    g = ggp.ForestGroup(d, [k % nd for k in range(shards)])
    for mode in ("strict", "fast"):
        g.set_mode(mode)
        for _ in range(5):
            ll = g.total_likelihood(P)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            ll = g.total_likelihood(P)
            ts.append(time.perf_counter() - t0)
        print("%d shards on %d device(s), %s: %.3f ms per evaluation (median of 20), loglik %.10g" % (shards, nd, mode, 1e3 * np.median(ts), ll))
    g.close()
