"""H2D bandwidth probe: pinned host -> device for the bench's series (probe, not product)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import gfp_gaussian_process_b200 as ggp
n = 12_600_000
a = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(3)]
d = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for x, y in zip(a, d): y.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"torch pinned copy 3 x {n * 8 / 1e6:.0f} MB: {dt * 1e3:.2f} ms = {3 * n * 8 / dt / 1e9:.1f} GB/s")
data = ggp.simulate_forest(10000, 6, noise_model="const", division_model="gauss", seed=1)
forest = ggp.Forest(data, device=0)
pin = [torch.from_numpy(x).pin_memory() for x in (data.time, data.log_length, data.fp)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    forest.upload_series(pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"upload_series alone ({data.time.size} ctp): {dt * 1e3:.2f} ms = {3 * data.time.size * 8 / dt / 1e9:.1f} GB/s")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    forest.upload_series(pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
    ll = ggp.total_likelihood(ggp.PARAMS_CONST_GAUSS, forest)
    dt = time.perf_counter() - t0
    print(f"upload + loglik: {dt * 1e3:.2f} ms")
