#!/usr/bin/env python
"""Per-role phase / barrier clocks of the cooperative likelihood step (profiling build only).
Build:  nvcc <flags of __graft_entry__.NVCC_FLAGS> -DGGP_PHASE_CLOCKS -o build/libggp_clk.so gfp_gaussian_process_b200/csrc/ggp_b200.cu
Run  :  GGP_B200_LIB=build/libggp_clk.so python tools/phase_clocks.py [trees]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import _lib

trees = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
data = ggp.simulate_forest(trees, 6, noise_model="const", division_model="gauss", seed=1)
forest = ggp.Forest(data, device=0)
lib = _lib.load()
for _ in range(3):
    ggp.total_likelihood(ggp.PARAMS_CONST_GAUSS, forest)
buf = (C.c_ulonglong * 40)()
lib.ggp_debug_phase_clocks(None, 1)
ll = ggp.total_likelihood(ggp.PARAMS_CONST_GAUSS, forest)
lib.ggp_debug_phase_clocks(buf, 0)
a = np.array(list(buf), dtype=np.float64).reshape(4, 10)
names = ["align", "ph0", "w0", "ph1", "w1", "ph2", "w2", "ph3", "w3", "-"]
tot = a.sum(axis=1)
print("loglik", ll)
print("share of a warp's time per slot (%), by role")
print("role " + " ".join(f"{n:>6}" for n in names[:9]))
for r in range(4):
    print(f"{r:4d} " + " ".join(f"{100 * a[r, k] / tot[r]:6.1f}" for k in range(9)))
print("work only (clocks per role, relative to role 0 total work):")
w = a[:, [1, 3, 5, 7]]
for r in range(4):
    print(f"{r:4d} " + " ".join(f"{x / w[0].sum():6.3f}" for x in w[r]) + f"   sum {w[r].sum() / w[0].sum():.3f}")
