"""End-to-end runs of the gfp_gaussian command line (host/gfp_gaussian.cpp over libggp_b200.so) on a GPU, checked
against the oracle: the files a user of the reference gets (iterations, final, parameter file, scan, prediction,
joints) with the reference's layout, and the numbers inside them."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, example_data, max_rel
from oracle.oracle_py import Oracle
import gfp_gaussian_process_b200 as ggp

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "gfp_gaussian_process_b200", "bin", "gfp_gaussian")
NAMES = ["mean_lambda", "gamma_lambda", "var_lambda", "mean_q", "gamma_q", "var_q", "beta", "var_x", "var_g", "var_dx", "var_dg"]


def write_inputs(tmp, data, segment_col=False):
    csv, cfg = str(tmp / "forest.csv"), str(tmp / "cfg.txt")
    with open(csv, "w") as f:
        f.write("lane,cell_id,parent_id,time_min,length,gfp,segment\n")
        for c in range(data.n_cells):
            p = int(data.parent[c])
            for k in range(data.cell_offset[c], data.cell_offset[c + 1]):
                f.write(f"1,{c + 1},{p + 1 if p >= 0 else 0},{float(data.time[k])!r},{float(data.log_length[k])!r},{float(data.fp[k])!r},{int(data.segment[k])}\n")
    with open(cfg, "w") as f:
        f.write("time_col = time_min\nlength_col = length\nlength_islog = true\nfp_col = gfp\ncell_tags = lane, cell_id\nparent_tags = lane, parent_id\n")
        if segment_col:
            f.write("segment_col = segment\n")
    return csv, cfg


def write_params(path, values, free=(), bound=()):
    with open(path, "w") as f:
        for i, (n, v) in enumerate(zip(NAMES, values)):
            if i in free:
                f.write(f"{n} = {float(v)!r}, {float(v) * 0.1!r}\n")
            elif i in bound:
                f.write(f"{n} = {float(v)!r}, {float(v) * 0.05!r}, {float(v) * 0.8!r}, {float(v) * 1.2!r}\n")
            else:
                f.write(f"{n} = {float(v)!r}\n")
    return str(path)


def run(args, expect=0):
    r = subprocess.run([CLI] + args, capture_output=True, text=True)
    assert r.returncode == expect, r.stdout + r.stderr
    return r


def g6(v):
    return "%g" % v   # default ostream formatting of a double


def read_table(path, header_prefix):
    rows, on = [], False
    for line in open(path):
        line = line.rstrip("\n")
        if on and line:
            rows.append(line.split(","))
        if line.startswith(header_prefix):
            on = True
    return rows


def test_prediction_file_matches_oracle_at_file_precision(tmp_path, golden_dir):
    data, z = example_data(golden_dir)
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", z["params"])
    out = str(tmp_path / "out")
    run(["-i", csv, "-b", pf, "-c", cfg, "-p", "-o", out])
    assert os.path.exists(os.path.join(out, "forest_success.log"))
    rows = read_table(os.path.join(out, "forest_f_b_prediction.csv"), "cell_id,parent_id,time")
    assert len(rows) == data.n_ctp
    pr = Oracle(data).predictions([z["params"]])["prediction"]
    iu = [(m, n) for m in range(4) for n in range(m, 4)]
    for k in list(range(0, data.n_ctp, 97)) + [data.n_ctp - 1]:
        want = [g6(v) for v in pr[0][k]] + [g6(pr[1][k][m][n]) for m, n in iu]
        assert rows[k][5:] == want, (k, rows[k][5:], want)
    assert rows[0][0] == "1.1" and rows[0][1] == "1.0"   # ids composed from the tags


def test_minimization_scan_and_error_bars_replay_on_the_oracle(tmp_path):
    """-m -s on a small forest: every line of the iterations / scan files is one evaluation; re-evaluating the logged
    parameter vectors on the oracle in the same order (history-dependent) reproduces the logged values"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    data = ggp.simulate_forest(8, 3, noise_model="scaled", division_model="binomial", seed=21)
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", P * np.array([1.1, 1, 1, 0.9, 1, 1, 1, 1, 1, 1, 1]), free=(0, 3), bound=(8,))
    out = str(tmp_path / "out")
    run(["-i", csv, "-b", pf, "-c", cfg, "-m", "-s", "-t", "1e-2", "-o", out])
    base = os.path.join(out, "forest_f03_b8")
    it = read_table(base + "_iterations.csv", "iteration,")
    assert [r[0] for r in it] == [str(i + 1) for i in range(len(it))] and len(it) > 8
    o = Oracle(data)
    o.reset()
    logged = np.array([[float(x) for x in r[1:]] for r in it])
    for row in logged:
        assert abs(o.total_loglik(row[:11], fresh=False) - row[11]) <= 1e-10 * abs(row[11])
    assert logged[:, 11].max() >= logged[0, 11]
    final = open(base + "_final.csv").read()
    assert "errors^2:\nepsilon,mean_lambda,mean_q,var_g\n0.05," in final and "total_log_likelihoood," in final
    assert "optimization_algorithm,LN_NELDERMEAD" in final and "search_space,log" in final
    ll_max = float(final.split("total_log_likelihoood,")[1].split("\n")[0])
    assert abs(ll_max - logged[:, 11].max()) <= 1e-9 * abs(ll_max)
    pfile = open(base + "_parameter_file.txt").read().split("\n")
    assert pfile[1].startswith("mean_lambda = ") and len(pfile) == 13
    # the scan runs on a fresh copy of the cells with the initial values (params are passed by value to the modes)
    sc = read_table(os.path.join(out, "forest_scan_var_g.csv"), "iteration,")
    v0 = P[8]
    grid = ggp.arange(v0 * 0.8, v0 * 1.2, v0 * 0.05)
    assert len(sc) == len(grid) and np.array_equal([float(r[9]) for r in sc], grid)
    o2 = Oracle(data)
    o2.reset()
    for r in sc:
        row = [float(x) for x in r[1:]]
        assert abs(o2.total_loglik(row[:11], fresh=False) - row[11]) <= 1e-10 * abs(row[11])


def test_fresh_mode_speculative_search_reaches_the_same_optimum(tmp_path):
    P = ggp.PARAMS_CONST_GAUSS
    data = ggp.simulate_forest(8, 3, seed=22)
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", P * np.array([1.2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]), free=(0,))
    res = []
    for extra, name in (([], "a"), (["--fresh"], "b")):
        out = str(tmp_path / name)
        run(["-i", csv, "-b", pf, "-c", cfg, "-m", "-t", "1e-6", "-noise", "const", "-div", "gauss", "-o", out] + extra)
        it = read_table(os.path.join(out, "forest_f0_b_iterations.csv"), "iteration,")
        res.append(np.array([[float(x) for x in r[1:]] for r in it]))
    best = [r[np.argmax(r[:, 11])] for r in res]
    # the two objectives differ by the reference's history dependence (stale root off-diagonals, ~1e-5 relative)
    assert abs(best[0][0] - best[1][0]) < 1e-3 * best[0][0] and abs(best[0][11] - best[1][11]) < 1e-4 * abs(best[0][11])
    o = Oracle(data)
    for row in res[1][:5]:
        assert abs(o.total_loglik(row[:11]) - row[11]) <= 1e-10 * abs(row[11])   # fresh: a pure function of the parameters


def test_fast_option_stays_inside_the_gate_on_every_logged_evaluation(tmp_path):
    """--fast: -m and -s with the fast likelihood arithmetic; every evaluation the command line logged is within 1e-10 of
    the oracle's value for the same parameters, and the predictions written afterwards are still the strict ones"""
    P = ggp.PARAMS_CONST_GAUSS
    data = ggp.simulate_forest(8, 3, seed=22)
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", P * np.array([1.2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]), free=(0,))
    out = str(tmp_path / "f")
    run(["-i", csv, "-b", pf, "-c", cfg, "-m", "-p", "-t", "1e-6", "-noise", "const", "-div", "gauss", "-o", out, "--fast"])
    it = read_table(os.path.join(out, "forest_f0_b_iterations.csv"), "iteration,")
    rows = np.array([[float(x) for x in r[1:]] for r in it])
    o = Oracle(data)
    for row in rows[:8]:
        assert abs(o.total_loglik(row[:11]) - row[11]) <= 1e-10 * abs(row[11])
    out2 = str(tmp_path / "s")
    run(["-i", csv, "-b", pf, "-c", cfg, "-m", "-p", "-t", "1e-6", "-noise", "const", "-div", "gauss", "-o", out2, "--fresh"])
    rows2 = np.array([[float(x) for x in r[1:]] for r in read_table(os.path.join(out2, "forest_f0_b_iterations.csv"), "iteration,")])
    assert abs(rows[np.argmax(rows[:, 11])][0] - rows2[np.argmax(rows2[:, 11])][0]) < 1e-6 * abs(rows2[0][0])


def test_joints_files_dense_and_sparse(tmp_path):
    P = ggp.PARAMS_SCALED_BINOMIAL
    data = ggp.simulate_forest(2, 3, noise_model="scaled", division_model="binomial", seed=15, pts_range=(3, 5))
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", P)
    o = Oracle(data)
    o.predictions([P])
    n, row, col, rec = o.joints(1e-10, 100000)
    order = np.lexsort((col, row))
    row, col, rec = row[order], col[order], rec[order]
    out = str(tmp_path / "dense")
    run(["-i", csv, "-b", pf, "-c", cfg, "-j", "-o", out])
    lines = open(os.path.join(out, "forest_f_b_joints.csv")).read().split("\n")
    h = [i for i, l in enumerate(lines) if l.startswith("cell_id,parent_id,time")][0]
    M = data.n_ctp
    assert len(lines[h].split(",")) == 3 + 44 * M                       # header: 3 + one 44-field block per column
    body = lines[h + 1:h + 1 + M]
    assert all(len(b.split(",")) == 3 + 44 * M for b in body)
    k = 0
    for r in range(M):
        fields = body[r].split(",")[3:]
        for c in range(M):
            blk = fields[44 * c:44 * c + 44]
            if k < n and row[k] == r and col[k] == c:
                assert blk == [g6(v) for v in rec[k]], (r, c)
                k += 1
            else:
                assert blk == [""] * 44
    assert k == n
    out = str(tmp_path / "sparse")
    run(["-i", csv, "-b", pf, "-c", cfg, "-j", "--sparse_joints", "-o", out])
    sp = read_table(os.path.join(out, "forest_f_b_joints.csv"), "cell_id,parent_id,time")
    assert len(sp) == n and sp[0][5:] == [g6(v) for v in rec[0]]


def test_nan_is_reported_like_the_reference(tmp_path):
    data = ggp.simulate_forest(3, 2, seed=5)
    csv, cfg = write_inputs(tmp_path, data)
    bad = ggp.PARAMS_CONST_GAUSS.copy()
    bad[8] = -1e9   # negative var_g: log of a negative determinant
    pf = write_params(tmp_path / "p.txt", bad, bound=(0,))
    out = str(tmp_path / "out")
    r = run(["-i", csv, "-b", pf, "-c", cfg, "-s", "-noise", "const", "-div", "gauss", "-o", out], expect=1)
    assert "Likelihood is Nan" in r.stdout
    log = open(os.path.join(out, "forest_error.log")).read()
    assert "Log likelihood is Nan" in log and "Cell: 1.1, observation: 0" in log


def test_devices_option_shards_trees(tmp_path):
    """--devices a,b: trees sharded over several handles (here twice the same GPU), log-likelihoods added on the host,
    predictions gathered by ctp: same files as the single-device run (totals to rounding, predictions identical)"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    data = ggp.simulate_forest(9, 3, noise_model="scaled", division_model="binomial", seed=23, pts_range=(3, 8))
    csv, cfg = write_inputs(tmp_path, data)
    pf = write_params(tmp_path / "p.txt", P, bound=(3,))
    outs = []
    for extra, name in (([], "one"), (["--devices", "0,0,0"], "three")):
        out = str(tmp_path / name)
        run(["-i", csv, "-b", pf, "-c", cfg, "-s", "-p", "-j", "-o", out] + extra)
        sc = read_table(os.path.join(out, "forest_scan_mean_q.csv"), "iteration,")
        pr = read_table(os.path.join(out, "forest_f_b3_prediction.csv"), "cell_id,parent_id,time")
        jn = open(os.path.join(out, "forest_f_b3_joints.csv")).read()
        outs.append((np.array([[float(x) for x in r[1:]] for r in sc]), pr, jn))
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]
    assert max_rel(outs[1][0][:, 11], outs[0][0][:, 11]) < 1e-13 and np.array_equal(outs[1][0][:, :11], outs[0][0][:, :11])
    # the correlation functions: every shard reduces its own trees' joints on its device, the sums add (ggp_group_correlation_sums)
    dt = repr(float(data.time[1] - data.time[0]))
    tabs = []
    for extra, name in (([], "corr_one"), (["--devices", "0,0,0"], "corr_three")):
        out = str(tmp_path / name)
        run(["-i", csv, "-b", pf, "-c", cfg, "-p", "--correlation", dt, "--n_data", "15", "-o", out] + extra)
        log = open([os.path.join(out, f) for f in os.listdir(out) if f.endswith("_success.log")][0]).read()
        assert "lag bins on the device" in log
        lines = open(os.path.join(out, "forest_f_b3_correlations.csv")).read().strip().split("\n")
        tabs.append(np.array([[float(x) if x else np.nan for x in l.split(",")] for l in lines[1:]]))
    a, b = tabs
    assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[:, -1], b[:, -1])
    ok = ~np.isnan(a)
    assert np.allclose(a[ok], b[ok], rtol=1e-9, atol=1e-12)


def test_binary_forest_input_gives_the_same_files(tmp_path):
    """-i forest.ggpf (binary forest file) instead of the csv of the same data: identical scan, prediction and joints files"""
    from gfp_gaussian_process_b200 import io
    P = ggp.PARAMS_SCALED_BINOMIAL
    data = ggp.simulate_forest(7, 3, noise_model="scaled", division_model="binomial", seed=29, pts_range=(3, 8))
    csv, cfg = write_inputs(tmp_path, data)
    binf = str(tmp_path / "forest.ggpf")
    io.write_forest_binary(binf, data, ["1.%d" % (c + 1) for c in range(data.n_cells)],
                           ["1.%d" % (int(p) + 1 if p >= 0 else 0) for p in data.parent])
    pf = write_params(tmp_path / "p.txt", P, bound=(3,))
    outs = []
    for infile, name in ((csv, "csv"), (binf, "bin")):
        out = str(tmp_path / name)
        run(["-i", infile, "-b", pf, "-c", cfg, "-s", "-p", "-j", "--sparse_joints", "-o", out])
        outs.append([open(os.path.join(out, "forest_" + f)).read() for f in ("scan_mean_q.csv", "f_b3_prediction.csv", "f_b3_joints.csv")])
    assert outs[0] == outs[1]
