#!/usr/bin/env python
"""GPU timing of the likelihood kernels on BASELINE configs[1]: strict and fast (several node counts), device time of
ggp_loglik_device (CUDA events inside the library), and the gate figure of each fast variant against the strict result."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import _lib
import torch

trees = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
modes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 5, 6, 8, 10]   # 0 = strict
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = _lib.load()
P = ggp.PARAMS_CONST_GAUSS
d = ggp.simulate_forest(trees, 6, seed=20261018)
f = ggp.Forest(d)
dp = torch.tensor(P, dtype=torch.float64, device="cuda").reshape(1, 11)
out = torch.zeros(1, dtype=torch.float64, device="cuda")
def run(reps=reps):
    ms = []
    for _ in range(reps + 3):
        _lib.check(lib.ggp_loglik_device(f.handle, dp.data_ptr(), 1, out.data_ptr()))
        k = np.zeros(1)
        _lib.check(lib.ggp_sync_kernel_ms(f.handle, k.ctypes.data_as(_lib.c_double_p)))
        ms.append(float(k[0]))
    return float(np.median(ms[3:])), float(out.item())
ll_s = float("nan")
if 0 in modes:
    t, ll_s = run()
    print(f"strict: {t:.3f} ms  {d.n_ctp / t / 1e6:.2f} Gctp/s  loglik {ll_s!r}")
for n in [m for m in modes if m]:
    f.set_mode(n)
    try:
        t, ll = run()
        print(f"fast N={n}: {t:.3f} ms  {d.n_ctp / t / 1e6:.2f} Gctp/s  rel vs strict {abs(ll - ll_s) / abs(ll_s):.2e}")
    except Exception as e:
        print(f"fast N={n}: {e}")
f.close()
