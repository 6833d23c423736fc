"""N ranks copy 2 x 100.8 MB of pinned host memory to their GPUs AT THE SAME TIME (barrier before every copy): the host-side
ceiling of the end-to-end step at N GPUs.  torchrun --nproc-per-node N tools/r02_jobs/h2d_concurrent.py"""
import os, time
import numpy as np
import torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 12_599_120
pin = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
dev = [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(2)]
for p in pin:
    p.fill_(1.0)
res = {}
for mode in ("alone", "together"):
    ts = []
    for rep in range(8):
        dist.barrier(); torch.cuda.synchronize()
        if mode == "alone":     # ranks take turns
            for r in range(world):
                if r == rank:
                    t0 = time.perf_counter()
                    for d, p in zip(dev, pin): d.copy_(p, non_blocking=True)
                    torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                dist.barrier()
        else:
            t0 = time.perf_counter()
            for d, p in zip(dev, pin): d.copy_(p, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.median(ts[2:]))], dtype=torch.float64, device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    res[mode] = [float(x.item()) for x in allt]
if rank == 0:
    gb = 2 * 8 * n / 1e9
    for mode in ("alone", "together"):
        per = [gb / x for x in res[mode]]
        print("%s: per GPU %s GB/s; %s" % (mode, " ".join("%.1f" % v for v in per),
              "aggregate %.1f GB/s" % (world * gb / max(res[mode])) if mode == "together" else "one at a time"))
dist.destroy_process_group()
