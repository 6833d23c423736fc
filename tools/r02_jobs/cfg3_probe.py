"""configs[2] -p on one GPU: kernel time and end to end into pinned host memory (14 and 20 doubles per point)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import _lib
lib = _lib.load()
data = ggp.simulate_forest(15873, 6, noise_model="scaled", division_model="binomial", seed=20261018)
f = ggp.Forest(data)
P = np.ascontiguousarray(ggp.PARAMS_SCALED_BINOMIAL.reshape(1, 11))
pins = {k: torch.empty((data.n_ctp, 14), dtype=torch.float64).pin_memory().numpy() for k in ("forward", "backward", "prediction")}
for _ in range(4):
    t0 = time.perf_counter()
    ggp.prediction_upper14(f, P, out=pins)
    dt = time.perf_counter() - t0
    print("predict14 e2e %.1f ms (kernels %.2f ms), %.1f GB/s" % (dt * 1e3, f.last_kernel_ms, 3 * 14 * 8 * data.n_ctp / dt / 1e9))
print("checksum", float(pins["prediction"][::1009].sum()), float(pins["forward"][::1009].sum()), float(pins["backward"][::1009].sum()))
f.close()
