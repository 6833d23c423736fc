#!/usr/bin/env python
"""bench.py — throughput of the lineage-forest log-likelihood (FP64) on N B200s.

Workload (BASELINE.json configs[1]): synthetic forest of 10 000 lineage trees x 6 generations x ~20 points
per cell PER GPU (630 000 cells, ~12.6 M cell-timepoints), one log-likelihood evaluation per step,
"gauss" division, "const" noise, fresh mode.  Weak scaling: every rank simulates its own forest (seed + rank);
the init_cells_f/r population statistics are combined across ranks (all-reduce of 10 sums) and the scalar
log-likelihood is all-reduced every step (NCCL).  Inputs (353 MB/GPU) are larger than the 126 MB L2.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference ...                         the reference's CPU math on the host cores

One JSON line on stdout (rank 0).  `value` = cell-timepoints/s of the whole job with inputs resident in HBM,
`e2e` = the same through the host-buffer C-ABI call (pinned host series uploaded and the result read back
inside the timed region), `roofline` = FP64-pipe fraction of the likelihood kernel, `cpu_baseline` = the
oracle on one host core on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_ALG = 3700.0        # algorithmic FP64 flop per cell-timepoint per directional pass
B_ALG = 28.0          # algorithmic bytes per cell-timepoint: time, x, g (f64) + segment (i32)
# dram__bytes_read.sum + dram__bytes_write.sum of the likelihood kernel per cell-timepoint, from the ncu --set full capture
# of the largest generation's launch (profiles/r01_s5_loglik_coop_gen5.txt, 6 399 691 ctp)
DRAM_BYTES_PER_CTP_NCU = (204.285440e6 + 4.971776e6) / 6399691.0
METRIC = "cell-timepoints/s, FP64 log-likelihood evaluation (loglik evals/s in config)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trees", type=int, default=10000)
    ap.add_argument("--generations", type=int, default=6)
    ap.add_argument("--cpu-sample-trees", type=int, default=1500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": "configs[1]: synthetic forest %d trees x %d generations x ~20 pts/cell per GPU, single "
                        "log-likelihood eval, gauss division, const noise, fresh mode" % (args.trees, args.generations),
            "trees_per_gpu": args.trees, "generations": args.generations, "n_vec": 1, "parallelism": "trees sharded x%d" % world,
            "l2": "inputs (28 B/ctp, ~353 MB per GPU) larger than the 126 MB L2"}


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's filter loop around the reference's own mean_cov_model.h +
# Faddeeva.cc (oracle/_ref) where that build exists, else the oracle port
# ---------------------------------------------------------------------------------------------------
def _cpu_worker(q_in, q_out, use_ref):
    from oracle import oracle_py
    if use_ref:
        oracle_py.use_reference_math(True)
    import gfp_gaussian_process_b200 as ggp
    o = None
    while True:
        msg = q_in.get()
        if msg is None:
            break
        if msg[0] == "load":
            z = np.load(msg[1])
            data = ggp.LineageData(cell_offset=z["cell_offset"], parent=z["parent"], time=z["time"], log_length=z["log_length"],
                                   fp=z["fp"], noise_model="const", division_model="gauss", init_f=z["init_f"], init_r=z["init_r"])
            o = oracle_py.Oracle(data)
            q_out.put(("loaded", data.n_ctp))
        else:
            t = time.perf_counter()
            ll = o.total_loglik(msg[1])
            q_out.put((ll, time.perf_counter() - t))


def cpu_pool_bench(data, params, n_proc, trees_per_proc, steps, warmup, use_ref):
    """P processes, each evaluating its own shard of `trees_per_proc` trees per step; returns (ctp/s, ctp per step, loglik)"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    roots = data.roots()
    init_f, init_r = data.init_stats()
    procs, files, total = [], [], 0
    tmp = tempfile.mkdtemp(prefix="ggp_bench_")
    for r in range(n_proc):
        sub, _, _ = data.subset(roots[r * trees_per_proc:(r + 1) * trees_per_proc])
        fn = os.path.join(tmp, f"shard{r}.npz")
        np.savez(fn, cell_offset=sub.cell_offset, parent=sub.parent, time=sub.time, log_length=sub.log_length, fp=sub.fp,
                 init_f=init_f, init_r=init_r)
        files.append(fn)
        qi, qo = ctx.Queue(), ctx.Queue()
        p = ctx.Process(target=_cpu_worker, args=(qi, qo, use_ref), daemon=True)
        p.start()
        procs.append((p, qi, qo))
    for (p, qi, qo), fn in zip(procs, files):
        qi.put(("load", fn))
    for p, qi, qo in procs:
        total += qo.get()[1]
    times, ll = [], 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for p, qi, qo in procs:
            qi.put(("eval", params))
        res = [qo.get() for p, qi, qo in procs]
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        ll = sum(r[0] for r in res)
    for p, qi, qo in procs:
        qi.put(None)
    for p, qi, qo in procs:
        p.join(timeout=10)
    for fn in files:
        os.remove(fn)
    os.rmdir(tmp)
    return total * len(times) / sum(times), total, ll, sum(times) / len(times)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import gfp_gaussian_process_b200 as ggp
    from oracle import oracle_py
    use_ref = oracle_py.use_reference_math(True)
    oracle_py.use_reference_math(False)
    cores = min(host_cores(), 64)
    trees_per_proc = 120          # ~150 k ctp, ~1 s per process per step
    data = ggp.simulate_forest(cores * trees_per_proc, args.generations, seed=20261018)
    steps, warmup = min(args.steps, 20), min(args.warmup, 3)   # ~1 s per step on the sample below
    v, ctp, ll, t = cpu_pool_bench(data, ggp.PARAMS_CONST_GAUSS, cores, trees_per_proc, steps, warmup, use_ref)
    kind = "reference" if use_ref else "port"
    sample = ("%d trees x %d generations (%d ctp) of the configs[1] forest per step, %d processes x %d trees; %s"
              % (cores * trees_per_proc, args.generations, ctp, cores, trees_per_proc,
                 "reference mean_cov_model.h + Faddeeva.cc compiled unmodified (oracle/_ref), filter loop = oracle restatement "
                 "(Eigen is not installable here)" if use_ref else "oracle port (oracle/ggp_oracle.cpp)"))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "ctp/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": "ctp/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "ctp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "loglik_evals_per_s_full_forest": v / (args.trees * (2 ** args.generations - 1) * 20.0), "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None
        self.f = None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, power, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "power_w_max": float(max(power)), "samples": len(sm)}
        return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gfp_gaussian_process_b200 as ggp
    from gfp_gaussian_process_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: native libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    P = ggp.PARAMS_CONST_GAUSS

    data = ggp.simulate_forest(args.trees, args.generations, seed=20261018 + rank)
    # population statistics over ALL ranks' cells (moma_input.h:675-735): all-reduce of counts and sums
    n = np.diff(data.cell_offset)
    sel = n > 1
    sums = []
    for idx in (data.cell_offset[:-1][sel], data.cell_offset[1:][sel] - 1):
        x, g = data.log_length[idx], data.fp[idx]
        sums += [float(len(idx)), x.sum(), g.sum(), (x * x).sum(), (g * g).sum()]
    st = torch.tensor(sums, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(st)
    st = st.cpu().numpy()
    for k, name in ((0, "init_f"), (5, "init_r")):
        c, sx, sg, sxx, sgg = st[k:k + 5]
        setattr(data, name, np.array([sx / c, sg / c, sxx / c - (sx / c) ** 2, sgg / c - (sg / c) ** 2]))

    forest = ggp.Forest(data, device=local)
    stream = torch.cuda.current_stream()
    forest.set_stream(stream.cuda_stream)
    n_ctp = forest.n_ctp
    d_params = torch.tensor(P, dtype=torch.float64, device="cuda").reshape(1, 11)
    d_out = torch.zeros(1, dtype=torch.float64, device="cuda")
    total = torch.zeros(1, dtype=torch.float64, device="cuda")

    def step():
        _lib.check(lib.ggp_loglik_device(forest.handle, d_params.data_ptr(), 1, d_out.data_ptr()))
        total.copy_(d_out)
        if world > 1:
            dist.all_reduce(total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = 0.0
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = forest.last_launch_count * args.steps
    if rank == 0:
        time.sleep(0.2)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ll_total = float(total.item())

    # kernel-only time of the likelihood launches (events inside the library, same stream), separate loop so that the
    # event sync does not sit in the timed region above
    for _ in range(args.steps):
        _lib.check(lib.ggp_loglik_device(forest.handle, d_params.data_ptr(), 1, d_out.data_ptr()))
        k = np.zeros(1)
        _lib.check(lib.ggp_sync_kernel_ms(forest.handle, k.ctypes.data_as(_lib.c_double_p)))
        kern_ms += float(k[0])
    kern_ms /= args.steps

    # end to end through the host-buffer C-ABI call: pinned host series -> device, host params in, host result out
    pin = [torch.from_numpy(a).pin_memory() for a in (data.time, data.log_length, data.fp)]
    torch.cuda.synchronize()

    def e2e_step():
        forest.upload_series(pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
        return ggp.total_likelihood(P, forest)

    for _ in range(2):
        ll_e2e = e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ll_e2e = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    lle = torch.tensor([ll_e2e], dtype=torch.float64, device="cuda")
    ctp_all = torch.tensor([float(n_ctp)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(lle)
        dist.all_reduce(ctp_all)
    e2e_s = float(te.item())
    ctp_total = float(ctp_all.item())

    if rank == 0:
        peak = np.zeros(1)
        _lib.check(lib.ggp_fp64_peak(local, peak.ctypes.data_as(_lib.c_double_p)))
        fp64_peak = float(peak[0])
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        ms_step = ms / args.steps
        achieved_tf = n_ctp * F_ALG / (kern_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": ctp_total / (ms_step * 1e-3), "unit": "ctp/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "loglik_evals_per_s": 1e3 / ms_step, "loglik": ll_total, "loglik_e2e": float(lle.item()),
            "ctp_total": ctp_total, "cells_per_gpu": forest.n_cells,
            "clocks": clocks,
            "e2e": {"value": ctp_total / e2e_s, "unit": "ctp/s", "h2d_bytes_per_step": int(3 * 8 * n_ctp + 88),
                    "d2h_bytes_per_step": 16, "ms_per_step": e2e_s * 1e3,
                    "what": "ggp_forest_upload_series (time, log_length, fp from pinned host memory) + ggp_loglik (host params in, "
                            "host log-likelihood and NaN record out), wall clock around the host calls"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                         "traffic": DRAM_BYTES_PER_CTP_NCU * n_ctp, "traffic_unit": "bytes per step (sum over the step's launches; ncu bytes/ctp of "
                                                                                    "the largest launch x ctp per step)",
                         "kernel": "ggp_loglik_coop_kernel (%d launches/step, one per generation)" % forest.n_generations,
                         "kernel_ms_per_step": kern_ms, "flop_per_ctp": F_ALG,
                         "peak_source": "DFMA micro-benchmark run in this process (ggp_fp64_peak); MEASURED_PEAKS.json holds no FP64 figure",
                         "hbm": {"achieved": n_ctp * B_ALG / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": n_ctp * B_ALG / (kern_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}},
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle_py
            use_ref = oracle_py.use_reference_math(True)
            sub, _, _ = data.subset(data.roots()[:args.cpu_sample_trees])
            sub.init_f, sub.init_r = data.init_f, data.init_r
            o = oracle_py.Oracle(sub)
            t0 = time.perf_counter()
            ll_cpu = o.total_loglik(P)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sub.n_ctp / dt, "unit": "ctp/s", "cores": 1, "kind": "reference" if use_ref else "port",
                                    "sample": "first %d trees (%d ctp) of the same forest, one evaluation, %.1f s; %s" % (
                                        args.cpu_sample_trees, sub.n_ctp, dt,
                                        "reference mean_cov_model.h + Faddeeva.cc (oracle/_ref) inside the oracle's filter loop" if use_ref
                                        else "oracle port"), "loglik": ll_cpu}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    forest.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
