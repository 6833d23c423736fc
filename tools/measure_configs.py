#!/usr/bin/env python
"""Measures the BASELINE.json configurations other than the bench's (configs[1]) on one GPU and prints one JSON line
each; the numbers quoted in DESIGN.md come from here.  usage: python tools/measure_configs.py [cfg1 cfg3 cfg4 cfg5]"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gfp_gaussian_process_b200 as ggp  # noqa: E402
from gfp_gaussian_process_b200 import _lib  # noqa: E402

which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4", "cfg5"]


def out(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts))


if "cfg1" in which:   # the example data set (one tree, 77 cells, 22 065 points): per-evaluation latency and the command line
    from conftest import example_data
    data, z = example_data(os.path.join(ROOT, "tests", "golden"))
    f = ggp.Forest(data)
    P = z["params"]
    t1 = timed(lambda: ggp.total_likelihood(P, f))
    t64 = timed(lambda: ggp.total_likelihood(np.tile(P, (64, 1)), f))
    t400 = timed(lambda: ggp.total_likelihood(np.tile(P, (400, 1)), f))
    tp = timed(lambda: ggp.prediction_forward_backward(f, [P]))
    f.set_mode("fast")
    f1 = timed(lambda: ggp.total_likelihood(P, f))
    f400 = timed(lambda: ggp.total_likelihood(np.tile(P, (400, 1)), f))
    out(config="cfg1 example data set, library calls", n_ctp=int(data.n_ctp), generations=f.n_generations, ms_loglik_1=t1 * 1e3,
        ms_loglik_64=t64 * 1e3, ms_loglik_400=t400 * 1e3, ms_predict=tp * 1e3, ms_fast_loglik_1=f1 * 1e3, ms_fast_loglik_400=f400 * 1e3,
        fast_nodes=int(f.last_fast_nodes), fast_strict_reruns=int(f.last_strict_reruns))
    f.close()
    from test_gpu_cli import write_inputs, write_params, CLI
    import pathlib
    tmp = pathlib.Path(tempfile.mkdtemp())
    csv, cfg = write_inputs(tmp, data)
    pf = str(tmp / "p.txt")
    with open(pf, "w") as fh:   # the example's parameter file: everything free except beta
        for i, (n, v) in enumerate(zip(ggp.PARAM_NAMES, P)):
            fh.write(f"{n} = {float(v)!r}\n" if i == 6 else f"{n} = {float(v)!r}, {float(v) * 0.1!r}\n")
    for extra, name in ((["--fast"], "fast"), (["--fresh"], "fresh"), ([], "carry")):
        t = time.perf_counter()
        r = subprocess.run(["timeout", "900", CLI, "-i", csv, "-b", pf, "-c", cfg, "-m", "-p", "-o", str(tmp / name)] + extra,
                           capture_output=True, text=True)
        dt = time.perf_counter() - t
        log = ""
        for fn in os.listdir(tmp / name) if os.path.exists(tmp / name) else []:
            if fn.endswith(".log"):
                log = open(tmp / name / fn).read()
        stopped = [l for l in log.split("\n") if l.startswith("Stopped") or l.startswith("Found maximum")]
        out(config="cfg1 example data set, gfp_gaussian -m -p " + " ".join(extra), rc=r.returncode, wall_s=dt, log=stopped)

if "cfg2" in which:   # the command line on a configs[1]-size data set: binary forest file in, 1-d scan of mean_q (17 evaluations in one launch)
    from gfp_gaussian_process_b200 import io
    from test_gpu_cli import write_params, CLI
    import pathlib
    tmp = pathlib.Path(tempfile.mkdtemp())
    data = ggp.simulate_forest(10000, 6, seed=20261018)
    t = time.perf_counter()
    io.write_forest_binary(str(tmp / "forest.ggpf"), data)
    t_write = time.perf_counter() - t
    open(tmp / "cfg.txt", "w").write("fp_auto = 0\n")
    P = ggp.PARAMS_CONST_GAUSS
    pf = write_params(tmp / "p.txt", P, bound=(3,))
    for extra, name in ((["--fast"], "fast"), (["--fresh"], "fresh")):
        t = time.perf_counter()
        r = subprocess.run(["timeout", "900", CLI, "-i", str(tmp / "forest.ggpf"), "-b", pf, "-c", str(tmp / "cfg.txt"), "-s", "-noise", "const",
                            "-div", "gauss", "-o", str(tmp / name)] + extra, capture_output=True, text=True)
        dt = time.perf_counter() - t
        n_lines = sum(1 for _ in open(tmp / name / "forest_scan_mean_q.csv")) if r.returncode == 0 else -1
        out(config="cfg2 command line, binary forest file (%d MB) -s mean_q %s" % (os.path.getsize(tmp / "forest.ggpf") >> 20, " ".join(extra)),
            rc=r.returncode, wall_s=dt, scan_file_lines=n_lines, write_binary_s=t_write, n_ctp=int(data.n_ctp), tail=r.stdout[-300:])

if "cfg3" in which:   # -p on a 1M-cell forest, binomial + scaled (one GPU's worth; 8 GPUs shard the trees)
    data = ggp.simulate_forest(15873, 6, noise_model="scaled", division_model="binomial", seed=20261018)
    f = ggp.Forest(data)
    lib = _lib.load()
    P = np.ascontiguousarray(ggp.PARAMS_SCALED_BINOMIAL.reshape(1, 11))
    ms = []
    for _ in range(3):
        _lib.check(lib.ggp_predict(f.handle, P.ctypes.data_as(_lib.c_double_p), 1, None, None, None))
        ms.append(f.last_kernel_ms)
    out(config="cfg3 -p forward+backward+combine, 1M cells, scaled/binomial, 1 GPU", n_cells=int(data.n_cells), n_ctp=int(data.n_ctp),
        kernel_ms=float(np.median(ms)), ctp_per_s=data.n_ctp / (np.median(ms) * 1e-3), flop_per_ctp=8900,
        tflops=data.n_ctp * 8900 / (np.median(ms) * 1e-3) / 1e12)
    f.close()

if "cfg4" in which:   # scan: 4096 parameter vectors over the 10k-tree forest in one call
    data = ggp.simulate_forest(10000, 6, seed=20261018)
    f = ggp.Forest(data)
    P0 = ggp.PARAMS_CONST_GAUSS
    vecs = []
    for i in range(11):
        for s in np.linspace(0.8, 1.2, 373 if i < 4 else 372):
            v = P0.copy()
            v[i] *= s
            vecs.append(v)
    vecs = np.array(vecs[:4096])
    t = time.perf_counter()
    ll = ggp.total_likelihood(vecs, f, raise_on_nan=False)
    dt = time.perf_counter() - t
    out(config="cfg4 scan 4096 vectors x 10k-tree forest, one ggp_loglik call, fresh", n_vec=len(vecs), n_ctp=int(data.n_ctp), wall_s=dt,
        kernel_ms=f.last_kernel_ms, evals_per_s=len(vecs) / dt, ctp_per_s=len(vecs) * data.n_ctp / dt,
        tflops=len(vecs) * data.n_ctp * 3700 / dt / 1e12, finite=int(np.isfinite(ll).sum()))
    f.close()

if "cfg5" in which:   # -j on a 100k-cell forest with two segments (sparse output, counted in row blocks)
    data = ggp.simulate_forest(1587, 6, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
    f = ggp.Forest(data)
    P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    t = time.perf_counter()
    ggp.prediction_forward_backward(f, P, forward=False, backward=False, combined=False)
    tp = time.perf_counter() - t
    import ctypes as C
    lib = _lib.load()
    cnt = C.c_int64(0)
    rows = min(data.n_ctp, 200000)
    t = time.perf_counter()
    _lib.check(lib.ggp_joints(f.handle, P.ctypes.data_as(_lib.c_double_p), 2, C.c_double(1e-10), 0, rows, 0, C.byref(cnt), None, None, None))
    dt = time.perf_counter() - t
    first_ms = f.last_kernel_ms      # per-point preparation (all points, cached on the handle) + walk
    walk_ms = []
    for _ in range(3):
        _lib.check(lib.ggp_joints(f.handle, P.ctypes.data_as(_lib.c_double_p), 2, C.c_double(1e-10), 0, rows, 0, C.byref(cnt), None, None, None))
        walk_ms.append(f.last_kernel_ms)
    # with records: walk + device sort into (row, col) order + copy into the caller's (pageable) arrays
    capn = int(cnt.value)
    hrow = np.empty(capn, dtype=np.int64); hcol = np.empty(capn, dtype=np.int64); hrec = np.empty((capn, 44))
    hrec[:] = 0.0   # touch the pages
    rec_s = []
    for _ in range(2):
        t = time.perf_counter()
        _lib.check(lib.ggp_joints(f.handle, P.ctypes.data_as(_lib.c_double_p), 2, C.c_double(1e-10), 0, rows, capn, C.byref(cnt),
                                  hrow.ctypes.data_as(_lib.c_int64_p), hcol.ctypes.data_as(_lib.c_int64_p), hrec.ctypes.data_as(_lib.c_double_p)))
        rec_s.append(time.perf_counter() - t)
    assert (np.diff(hrow) >= 0).all()
    out(config="cfg5 -j with records: first %d start points, %d records (%.2f GB) into host arrays in (row, col) order" % (rows, capn, capn * 368 / 1e9),
        wall_s=float(np.min(rec_s)), records_per_s=capn / float(np.min(rec_s)))
    del hrow, hcol, hrec
    all_ms = []
    call = C.c_int64(0)
    for _ in range(2):
        _lib.check(lib.ggp_joints(f.handle, P.ctypes.data_as(_lib.c_double_p), 2, C.c_double(1e-10), 0, int(data.n_ctp), 0, C.byref(call), None, None, None))
        all_ms.append(f.last_kernel_ms)
    out(config="cfg5 -j 100k cells, 2 segments, tol 1e-10, every start point, count only", n_ctp=int(data.n_ctp), joints=int(call.value),
        walk_ms=float(np.min(all_ms)), walk_joints_per_s=call.value / (np.min(all_ms) * 1e-3))
    out(config="cfg5 -j 100k cells, 2 segments, tol 1e-10 (first %d start points, count only)" % rows, n_cells=int(data.n_cells),
        n_ctp=int(data.n_ctp), predict_s=tp, joints=int(cnt.value), joints_per_start=cnt.value / rows, wall_s=dt, kernel_ms=first_ms,
        walk_ms=float(np.median(walk_ms)), joints_per_s=cnt.value / dt, walk_joints_per_s=cnt.value / (np.median(walk_ms) * 1e-3))
    f.close()
