"""e2e loop of the bench (fast mode) with the library's chunk timeline; prints ms per step"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gfp_gaussian_process_b200 as ggp
mode = sys.argv[1] if len(sys.argv) > 1 else "fast"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
P = ggp.PARAMS_CONST_GAUSS
data = ggp.simulate_forest(10000, 6, seed=20261018)
data.init_f, data.init_r = data.init_stats()
f = ggp.Forest(data, device=0)
f.set_mode(mode)
ggp.total_likelihood(P, f)
pin = [torch.from_numpy(a).pin_memory() for a in (data.log_length, data.fp)]
def step():
    f.upload_series(None, pin[0].data_ptr(), pin[1].data_ptr(), data.init_f, data.init_r)
    return ggp.total_likelihood(P, f)
for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    t0 = time.perf_counter()
    step()
    ts.append(time.perf_counter() - t0)
print("e2e %s chunks=%s: median %.3f ms  min %.3f ms  (copy-only floor %.3f ms at 55 GB/s)" % (
    mode, os.environ.get("GGP_B200_UPLOAD_CHUNKS", "6"), 1e3 * np.median(ts), 1e3 * min(ts), 16 * data.n_ctp / 55e9 * 1e3))
f.close()
