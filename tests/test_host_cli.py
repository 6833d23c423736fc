"""CPU tests of the gfp_gaussian command line's host side (host/*.hpp, host/gfp_gaussian.cpp): readers, segment
slicing, genealogy, parameter tables, file-name codes, the arange grid, the bounded Nelder-Mead and the binary's
behaviour without a GPU.  Numbers come from the GPU only: see tests/test_gpu_cli.py for the end-to-end runs."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gfp_gaussian_process_b200 import io as ggio

HC = os.path.join(ROOT, "tests", "hostcheck", "libhostcli.so")
CLI = os.path.join(ROOT, "gfp_gaussian_process_b200", "bin", "gfp_gaussian")


@pytest.fixture(scope="module")
def hc():
    lib = C.CDLL(HC)
    lib.hcli_load.restype = C.c_long
    lib.hcli_n_ctp.restype = C.c_long
    lib.hcli_cell_id.restype = C.c_char_p
    lib.hcli_text.restype = C.c_char_p
    return lib


def write_csv(path, data, cell_ids, parent_ids, segment_col=True, lane="1"):
    with open(path, "w") as f:
        f.write("lane,cell_id,parent_id,time_min,length,gfp,segment,keep\n")
        for c in range(data.n_cells):
            for k in range(data.cell_offset[c], data.cell_offset[c + 1]):
                f.write(f"{lane},{cell_ids[c]}.0,{parent_ids[c]},{float(data.time[k])!r},{float(np.exp(data.log_length[k]))!r},{float(data.fp[k])!r},"
                        f"{data.segment[k] if segment_col else 0},true\n")
            f.write(f"{lane},{cell_ids[c]}.0,{parent_ids[c]},0,1,1,0,false\n")   # filtered-out row


def load(hc, infile, config, segment=-1):
    n = hc.hcli_load(infile.encode(), config.encode(), segment)
    assert n >= 0, hc.hcli_text().decode()
    m = hc.hcli_n_ctp()
    off = np.zeros(n + 1, dtype=np.int64)
    par, d1, d2 = (np.zeros(n, dtype=np.int32) for _ in range(3))
    t, x, g = (np.zeros(m) for _ in range(3))
    seg = np.zeros(m, dtype=np.int32)
    hc.hcli_copy(off.ctypes.data_as(C.c_void_p), par.ctypes.data_as(C.c_void_p), d1.ctypes.data_as(C.c_void_p), d2.ctypes.data_as(C.c_void_p),
                 t.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p), seg.ctypes.data_as(C.c_void_p))
    ids = [hc.hcli_cell_id(c).decode() for c in range(n)]
    return dict(off=off, parent=par, d1=d1, d2=d2, time=t, x=x, g=g, seg=seg, ids=ids)


def test_example_dataset_loads_like_the_python_reader(hc, golden_dir):
    ex = "/root/reference/example_data_set"
    if not os.path.exists(ex):
        pytest.skip("reference tree not mounted")
    got = load(hc, ex + "/input.csv", ex + "/csv_config.txt")
    z = np.load(os.path.join(golden_dir, "example_forest.npz"))
    assert np.array_equal(got["off"], z["cell_offset"]) and np.array_equal(got["parent"], z["parent"])
    for k, name in (("time", "time"), ("x", "log_length"), ("g", "fp")):
        assert np.array_equal(got[k].view(np.uint64), z[name].view(np.uint64)), name   # same bits (same libm log)


def test_reader_filters_segments_and_genealogy(hc, tmp_path):
    from conftest import ragged_forest
    d = ragged_forest()
    ids = [str(100 + c) for c in range(d.n_cells)]
    pids = [ids[p] if p >= 0 else "7" for p in d.parent]
    csv, cfg = str(tmp_path / "in.csv"), str(tmp_path / "cfg.txt")
    write_csv(csv, d, ids, pids)
    open(cfg, "w").write("time_col = time_min\nlength_col = length\nfp_col = gfp\ncell_tags = lane, cell_id\nparent_tags = lane, parent_id\n"
                         "segment_col = segment\nfilter_col = keep\nfp_auto = 3\n")
    got = load(hc, csv, cfg)
    assert got["ids"] == ["1." + i for i in ids]                       # "100.0" -> "100", tags joined by '.'
    assert np.array_equal(got["off"], d.cell_offset) and np.array_equal(got["parent"], d.parent)
    assert np.array_equal(got["d1"], d.daughter1) and np.array_equal(got["d2"], d.daughter2)   # first / second child in file order
    assert np.array_equal(got["seg"], d.segment) and np.array_equal(got["time"], d.time)
    assert np.allclose(got["x"], d.log_length, rtol=1e-15, atol=1e-15)
    # one segment: cells without a point in it vanish, their daughters become roots (moma_input.h:580-620)
    s1 = load(hc, csv, cfg, segment=1)
    n1 = np.array([(d.segment[d.cell_offset[c]:d.cell_offset[c + 1]] == 1).sum() for c in range(d.n_cells)])
    kept = np.flatnonzero(n1 > 0)
    assert np.array_equal(np.diff(s1["off"]), n1[kept])
    remap = -np.ones(d.n_cells, dtype=np.int64)
    remap[kept] = np.arange(len(kept))
    assert np.array_equal(s1["parent"], [remap[p] if p >= 0 else -1 for p in d.parent[kept]])
    assert (s1["parent"] == -1).sum() >= (d.parent == -1).sum()


def test_reader_errors(hc, tmp_path):
    csv, cfg = str(tmp_path / "in.csv"), str(tmp_path / "cfg.txt")
    open(csv, "w").write("cell_id,parent_id,time,length,gfp\n1,0,0,2.0,10\n1,0,1,nan,11\n")
    open(cfg, "w").write("")
    assert hc.hcli_load(csv.encode(), cfg.encode(), -1) == -1 and b"Line no.3" in hc.hcli_text()
    open(cfg, "w").write("fp_col = missing\n")
    assert hc.hcli_load(csv.encode(), cfg.encode(), -1) == -1 and b"(fp_col) is not an column" in hc.hcli_text()
    open(csv, "w").write("cell_id,parent_id,time,length,gfp\n1,0,0,2,1\n2,1,1,2,1\n3,1,1,2,1\n4,1,1,2,1\n")
    open(cfg, "w").write("")
    assert hc.hcli_load(csv.encode(), cfg.encode(), -1) == -1 and b"Both daughter pointers are set" in hc.hcli_text()


def test_parameter_tables_and_codes(hc, tmp_path):
    pf = str(tmp_path / "p.txt")
    open(pf, "w").write("# comment\nmean_lambda = 0.01, 0.001\ngamma_lambda = 0.01, 1e-3, 1e-4, 0.1\nvar_lambda = 1e-7\nmean_q = 10, 1\n"
                        "gamma_q = 1e-2, 1e-3\nvar_q = 0.1, 0.01\nbeta = 5e-3\nvar_x = 1e-3, 1e-4\nvar_g = 5000, 50, 100, 1e5\n"
                        "var_dx = 1e-3\nvar_dg = 500, 50\n")
    assert hc.hcli_params(pf.encode(), None) == 0
    text = hc.hcli_text().decode()
    lines = text.split("\n")
    assert lines[0] == "no,name,type,init,step,lower_bound,upper_bound,final"
    assert lines[1] == "0,mean_lambda,free,0.01,0.001, , ,"
    assert lines[2] == "1,gamma_lambda,bound,0.01,0.001,0.0001,0.1,"
    assert lines[3] == "2,var_lambda,fixed,1e-07, , , ,"
    assert "CODE _f0345710_b18" in text   # free: 0 3 4 5 7 10 ; bound: 1 8
    ps = ggio.read_parameter_file(pf)
    assert [p.kind for p in ps][:3] == ["free", "bound", "fixed"]
    fin = (C.c_double * 11)(*[float(i + 1) for i in range(11)])
    assert hc.hcli_params(pf.encode(), fin) == 0
    assert hc.hcli_text().decode().split("\n")[2].endswith(",2")
    open(pf, "w").write("mean_lambda = 0.01, 0.001\n")
    assert hc.hcli_params(pf.encode(), None) == -1 and b"gamma_lambda not found" in hc.hcli_text()


def test_arange_accumulates(hc):
    n = C.c_int(0)
    out = (C.c_double * 64)()
    hc.hcli_arange(C.c_double(0.1), C.c_double(0.75), C.c_double(0.1), out, C.byref(n))
    v, ref, x = list(out)[:n.value], [], 0.1
    while x < 0.75:
        ref.append(x)
        x += 0.1
    assert v == ref and v != list(0.1 + 0.1 * np.arange(len(ref)))


def run_nm(hc, x0, lb, ub, step, ftol=1e-12, speculate=0):
    n = len(x0)
    arr = lambda v: (C.c_double * n)(*v)
    x = (C.c_double * n)()
    f, launches = C.c_double(0), C.c_int(0)
    ev = hc.hcli_neldermead(n, arr(x0), arr(lb), arr(ub), arr(step), C.c_double(ftol), speculate, x, C.byref(f), C.byref(launches))
    text = hc.hcli_text().decode().strip().split("\n")
    return np.array(list(x)), f.value, ev, launches.value, text[:-1], text[-1]


def test_nelder_mead_converges_respects_bounds_and_batches(hc):
    x, f, ev, launches, rec, why = run_nm(hc, [-1.2, 1.0, 0.5], [-5, -5, -5], [5, 5, 5], [0.1, 0.1, 0.1])
    assert why == "ftol reached" and np.allclose(x, 1.0, atol=1e-4) and f < 1e-9
    assert len(rec) == ev and launches < ev                     # initial simplex (and shrinks) are single launches
    # a fixed dimension (lb == ub) is eliminated; an active bound pins the optimum to the box
    x2, f2, *_ = run_nm(hc, [0.3, 0.5, 0.5], [-5, 0.5, -5], [0.6, 0.5, 5], [0.1, 1.0, 0.1])
    assert x2[1] == 0.5 and x2[0] <= 0.6 + 1e-15 and abs(x2[0] - 0.6) < 1e-3
    # a start on the upper bound steps inwards
    x3, f3, ev3, _, rec3, _ = run_nm(hc, [2.0, 2.0], [-3, -3], [2.0, 2.0], [0.5, 0.5])
    first = [list(map(float, r.split())) for r in rec3[:3]]
    assert first[0][:2] == [2.0, 2.0] and first[1][0] == 1.5 and first[2][1] == 1.5
    assert np.all(x3 <= 2.0) and f3 <= first[0][2]


def test_speculative_batching_replays_the_sequential_search(hc):
    a = run_nm(hc, [-1.2, 1.0, 0.5, 0.0], [-5] * 4, [5] * 4, [0.1] * 4)
    b = run_nm(hc, [-1.2, 1.0, 0.5, 0.0], [-5] * 4, [5] * 4, [0.1] * 4, speculate=1)
    assert a[4] == b[4] and np.array_equal(a[0], b[0]) and a[2] == b[2]   # same recorded evaluations, in the same order
    assert b[3] < 0.62 * a[3]                                             # ~1 launch per iteration instead of ~2


def test_deep_speculation_serves_several_iterations_per_launch(hc):
    """larger speculative batches hold the candidates of the following iterations for every way the current one can end: the
    recorded evaluations (points, values, order) stay those of the sequential search, the launches drop; a NaN the speculation
    meets is only reported where the sequential search evaluates that point itself"""
    hc.hcli_nm_nan_above.argtypes = [C.c_double]
    args = ([-1.2, 1.0, 0.5, 0.0], [-5] * 4, [5] * 4, [0.1] * 4)
    a = run_nm(hc, *args)
    one = run_nm(hc, *args, speculate=1)
    two = run_nm(hc, *args, speculate=28)
    three = run_nm(hc, *args, speculate=192)
    for b in (one, two, three):
        assert a[4] == b[4] and np.array_equal(a[0], b[0]) and a[2] == b[2] and a[5] == b[5]
    assert three[3] < two[3] < one[3] and two[3] < 0.62 * one[3] and three[3] < 0.45 * one[3]
    try:
        hc.hcli_nm_nan_above(1.02)       # the search overshoots the optimum at (1, 1, 1, 1): some evaluations are NaN
        s0 = run_nm(hc, *args)
        s3 = run_nm(hc, *args, speculate=192)
        assert any("nan" in r for r in s0[4]) and s0[4] == s3[4] and s0[2] == s3[2] and s0[5] == s3[5]
        assert np.array_equal(s0[0], s3[0], equal_nan=True)
    finally:
        hc.hcli_nm_nan_above(float("inf"))


def test_speculation_never_changes_the_search_on_random_problems(hc):
    """random starts, steps, boxes (some active, some dimensions fixed) and NaN regions: every speculation budget replays the
    sequential search evaluation by evaluation"""
    hc.hcli_nm_nan_above.argtypes = [C.c_double]
    rng = np.random.default_rng(4)
    try:
        for case in range(24):
            n = int(rng.integers(2, 6))
            x0 = rng.uniform(-2, 2, n)
            lb = x0 - rng.uniform(0.2, 4, n)
            ub = x0 + rng.uniform(0.2, 4, n)
            if case % 3 == 0:
                lb[0] = ub[0] = x0[0]            # a fixed dimension
            if case % 4 == 0:
                ub[-1] = min(ub[-1], 0.7)        # a bound the optimum (1, 1, ...) lies outside of
                x0[-1] = min(x0[-1], 0.7)
            step = rng.uniform(0.02, 0.5, n) * rng.choice([-1, 1], n)
            hc.hcli_nm_nan_above(float(rng.uniform(1.01, 1.5)) if case % 2 else float("inf"))
            ref = run_nm(hc, list(x0), list(lb), list(ub), list(step), ftol=1e-9)
            for budget in (1, 9, 40, 150):
                got = run_nm(hc, list(x0), list(lb), list(ub), list(step), ftol=1e-9, speculate=budget)
                assert got[4] == ref[4] and got[2] == ref[2] and got[5] == ref[5], (case, budget)
                assert np.array_equal(got[0], ref[0], equal_nan=True)
                if case % 2 == 0:                # (a NaN the search consumes is evaluated a second time, on its own)
                    assert got[3] <= ref[3], (case, budget)
    finally:
        hc.hcli_nm_nan_above(float("inf"))


def test_cli_without_gpu_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    csv = str(tmp_path / "in.csv")
    open(csv, "w").write("cell_id,parent_id,time,length,gfp\n1,0,0,2.0,10\n1,0,1,2.1,11\n2,1,2,1.1,5\n2,1,3,1.2,6\n")
    pf = str(tmp_path / "p.txt")
    open(pf, "w").write("\n".join(f"{n} = 1" for n in ggio.PARAM_NAMES) + "\n")
    r = subprocess.run([CLI, "-i", csv, "-b", pf, "-p", "-o", str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CUDA device" in r.stdout
    assert os.path.exists(tmp_path / "out" / "in_error.log") and not os.path.exists(tmp_path / "out" / "in_success.log")
    r = subprocess.run([CLI, "-i", csv, "-b", pf, "-noise", "poisson"], capture_output=True, text=True)
    assert r.returncode == 1 and "noise_model must be either" in r.stdout
    r = subprocess.run([CLI, "-h"], capture_output=True, text=True)
    assert r.returncode == 0 and "--parameter_bounds" in r.stdout


def test_python_reader_composes_ids_like_the_cpp_reader(hc, tmp_path):
    """remove_last_decimal (moma_input.h:327-347): only an all-zero fraction of a purely numeric tag is dropped; fractional
    and dotted tags stay, so '2.1' and '2.2' remain two cells.  Header tags are trimmed; bool spellings are strict."""
    csv, cfg = str(tmp_path / "in.csv"), str(tmp_path / "cfg.txt")
    rows = [("7.0", "0"), ("7.0", "0"), ("2.1", "7"), ("2.2", "7.00"), ("20150624.0.1.5", "2.1"), ("a.b", "2.2"), ("10.", "a.b")]
    with open(csv, "w") as f:
        f.write("cell_id , parent_id,time,length,gfp,keep\n")
        for k, (c, p) in enumerate(rows):
            f.write(f"{c},{p},{k},2.0,10,True\n")
    open(cfg, "w").write("filter_col = keep\n")
    got = load(hc, csv, cfg)
    data, ids = ggio.read_data(csv, ggio.read_csv_config(cfg))
    assert ids == got["ids"] == ["7", "2.1", "2.2", "20150624.0.1.5", "a.b", "10"]
    assert np.array_equal(data.parent, got["parent"]) and data.parent.tolist() == [-1, 0, 0, 1, 2, 4]
    assert np.array_equal(data.cell_offset, got["off"])
    with open(csv, "a") as f:
        f.write("11,10,9,2.0,10,yes\n")
    with pytest.raises(ValueError):
        ggio.read_data(csv, ggio.read_csv_config(cfg))
    assert hc.hcli_load(csv.encode(), cfg.encode(), -1) == -1


def test_binary_forest_file_round_trip(hc, tmp_path):
    """the binary forest file: written by the Python layer, read by the command line's loader and by the
    Python reader with the same arrays (bits), ids and genealogy as the csv of the same data; damaged files are refused"""
    from conftest import ragged_forest
    from gfp_gaussian_process_b200 import io
    d = ragged_forest()
    ids = ["1." + str(100 + c) for c in range(d.n_cells)]
    pids = [ids[p] if p >= 0 else "1.7" for p in d.parent]
    path = str(tmp_path / "forest.ggpf")
    io.write_forest_binary(path, d, ids, pids)
    cfg = str(tmp_path / "cfg.txt")
    open(cfg, "w").write("fp_auto = 3\n")
    got = load(hc, path, cfg)
    assert "binary file" in hc.hcli_text().decode() and got["ids"] == ids
    assert np.array_equal(got["off"], d.cell_offset) and np.array_equal(got["parent"], d.parent)
    assert np.array_equal(got["d1"], d.daughter1) and np.array_equal(got["d2"], d.daughter2) and np.array_equal(got["seg"], d.segment)
    for k, a in (("time", d.time), ("x", d.log_length), ("g", d.fp)):
        assert np.array_equal(got[k].view(np.uint64), np.asarray(a, dtype=np.float64).view(np.uint64))
    seg1 = load(hc, path, cfg, segment=1)                              # segment slicing works on it like on a csv
    assert seg1["time"].size == int((d.segment == 1).sum())
    back, cid, pid = io.read_forest_binary(path)
    assert cid == ids and pid == pids and np.array_equal(back.parent, d.parent) and np.array_equal(back.cell_offset, d.cell_offset)
    assert np.array_equal(back.fp.view(np.uint64), np.asarray(d.fp, dtype=np.float64).view(np.uint64))
    io.write_forest_binary(str(tmp_path / "default_ids.ggpf"), d)      # default ids: numbers from 1, 0 for "no parent"
    got2 = load(hc, str(tmp_path / "default_ids.ggpf"), cfg)
    assert got2["ids"][:3] == ["1", "2", "3"] and np.array_equal(got2["parent"], d.parent)
    raw = open(path, "rb").read()
    open(str(tmp_path / "cut.ggpf"), "wb").write(raw[:len(raw) // 2])
    assert hc.hcli_load(str(tmp_path / "cut.ggpf").encode(), cfg.encode(), -1) == -1 and "file ends" in hc.hcli_text().decode()
