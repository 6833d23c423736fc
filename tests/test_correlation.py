"""Correlation post-processing (host/ggp_correlation.hpp, gfp_gaussian --correlation / --correlation_files) against
golden vectors produced by the reference's own python_src/correlation_from_joint.py (tools/make_correlation_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from corr_files import CASES, write_case

CLI = os.path.join(ROOT, "gfp_gaussian_process_b200", "bin", "gfp_gaussian")
GRID_STEP = 2.0 / 9999   # the script's grid search over r


def read_table(path):
    lines = open(path).read().strip().split("\n")
    return lines[0].split(","), np.array([[float(x) for x in l.split(",")] for l in lines[1:]])


def check_against_reference(tab, ref_tab, ref_n):
    # our table has one more column (n_pairs) than corr_to_csv writes
    assert tab.shape == (ref_tab.shape[0], ref_tab.shape[1] + 1)
    assert np.array_equal(tab[:, -1], ref_n)                                 # pairs per lag: exact
    assert np.allclose(tab[:, 0], ref_tab[:, 0], rtol=0, atol=1e-12)         # dt grid
    naive, ref_naive = tab[:, 21:26], ref_tab[:, 21:26]
    ok = np.isfinite(ref_naive)
    assert np.array_equal(np.isfinite(naive), ok) and np.allclose(naive[ok], ref_naive[ok], rtol=1e-12, atol=1e-13)
    # grid-search maxima: the same grid point up to ties (numpy's vectorised log), error bars accordingly
    for k in range(1, 21, 2):
        scale = 1.0 if k >= 11 else np.maximum(np.abs(ref_tab[:, k]), 1e-300) / np.maximum(np.abs(ref_tab[:, k + 10]), 1e-12)
        fin = np.isfinite(ref_tab[:, k])
        assert np.all(np.abs(tab[fin, k] - ref_tab[fin, k]) <= 1.01 * GRID_STEP * (scale[fin] if np.ndim(scale) else scale) + 1e-15), k
    exact = np.isclose(tab[:, 11:21:2], ref_tab[:, 11:21:2], rtol=0, atol=1e-15)
    assert exact.mean() > 0.9                                                 # almost always the very same grid point
    err, ref_err = tab[:, 12:21:2][exact], ref_tab[:, 12:21:2][exact]
    assert np.allclose(err, ref_err, rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_correlation_from_files_matches_the_reference_script(name, tmp_path, golden_dir):
    z = np.load(os.path.join(golden_dir, "correlation_reference.npz"))
    case = CASES[name]
    jf, pf, dt = write_case(case, str(tmp_path))
    r = subprocess.run([CLI, "--correlation_files", jf, "--correlation", repr(dt), "--n_data", str(case["n_data"]), "-o", str(tmp_path / "log")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    head, tab = read_table(jf.replace("joints.csv", "correlations.csv"))
    assert head[0] == "dt" and head[1] == "cov_l(t+dt)l(t)" and head[25] == "corr_naive_c(t+dt)c(t)"
    check_against_reference(tab, z[name + "_table"], z[name + "_n"])
    if name + "_norm_table" in z.files:
        r = subprocess.run([CLI, "--correlation_files", jf, "--correlation", repr(dt), "--normalize_time", "-o", str(tmp_path / "log")],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        _, tab = read_table(jf.replace("joints.csv", "correlations.csv"))
        check_against_reference(tab, z[name + "_norm_table"], z[name + "_norm_n"])


@pytest.mark.gpu
def test_correlation_straight_from_the_gpu_joints(tmp_path, golden_dir):
    """--correlation: predictions + joints on the GPU, reduced without a joints file; against the reference script's result
    on the 6-digit files of the same forest (pair counts exact, moments to the files' precision)"""
    from test_gpu_cli import write_inputs, write_params
    import gfp_gaussian_process_b200 as ggp
    z = np.load(os.path.join(golden_dir, "correlation_reference.npz"))
    for name, case in CASES.items():
        P = ggp.PARAMS_CONST_GAUSS if case["noise"] == "const" else ggp.PARAMS_SCALED_BINOMIAL
        d = ggp.simulate_forest(case["trees"], case["gens"], params=P, noise_model=case["noise"], division_model=case["division"],
                                seed=case["seed"], pts_range=case["pts"])
        sub = tmp_path / name
        sub.mkdir()
        csv, cfg = write_inputs(sub, d)
        pf = write_params(sub / "p.txt", P)
        dt = float(np.min(np.diff(d.time[d.cell_offset[0]:d.cell_offset[1]]))) if d.cell_offset[1] > 1 else 1.0
        out = str(sub / "out")
        r = subprocess.run([CLI, "-i", csv, "-b", pf, "-c", cfg, "--correlation", repr(dt), "--n_data", str(case["n_data"]),
                            "-noise", case["noise"], "-div", case["division"], "-o", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert not os.path.exists(os.path.join(out, "forest_f_b_joints.csv"))
        _, tab = read_table(os.path.join(out, "forest_f_b_correlations.csv"))
        ref = z[name + "_table"]
        assert np.array_equal(tab[:, -1], z[name + "_n"])
        ok = np.isfinite(ref[:, 21:26])
        assert np.allclose(tab[:, 21:26][ok], ref[:, 21:26][ok], rtol=2e-4, atol=2e-5)   # inputs of the script carry 6 digits
        fin = np.isfinite(ref[:, 11:21:2])
        assert np.all(np.abs(tab[:, 11:21:2][fin] - ref[:, 11:21:2][fin]) <= 3 * GRID_STEP + 2e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("normalize", [False, True])
def test_lag_bins_accumulated_on_the_device_equal_the_host_reduction(tmp_path, normalize):
    """ggp_correlation_sums (the walk's records never leave the GPU) against the host reduction of the same sparse joints
    (host/ggp_correlation.hpp, itself pinned to the reference script above), both at full precision: pair counts exact,
    naive correlations to 1e-12, grid-search maxima on the same grid point with error bars to 1e-9; four
    generations deep, 40 lags"""
    from test_gpu_cli import write_inputs, write_params
    import gfp_gaussian_process_b200 as ggp
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(9, 4, params=P, noise_model="scaled", division_model="binomial", seed=61, pts_range=(5, 9))
    csv, cfg = write_inputs(tmp_path, d)
    pf = write_params(tmp_path / "p.txt", P)
    dt = float(d.time[1] - d.time[0])
    tabs = []
    for name, env in (("dev", {}), ("host", {"GGP_B200_CORR_HOST": "1"})):
        out = str(tmp_path / name)
        r = subprocess.run([CLI, "-i", csv, "-b", pf, "-c", cfg, "--correlation", repr(dt), "--n_data", "40", "-o", out] +
                           (["--normalize_time"] if normalize else []), capture_output=True, text=True, env={**os.environ, **env})
        assert r.returncode == 0, r.stdout + r.stderr
        log = open([os.path.join(out, f) for f in os.listdir(out) if f.endswith("_success.log")][0]).read()
        assert ("lag bins on the device" in log) == (name == "dev")
        tabs.append(read_table(os.path.join(out, "forest_f_b_correlations.csv"))[1])
    dev, host = tabs
    assert dev.shape == host.shape and np.array_equal(dev[:, -1], host[:, -1]) and dev[:, -1].sum() > 1000
    ok = np.isfinite(host[:, 21:26])
    assert np.array_equal(np.isfinite(dev[:, 21:26]), ok)
    zok = ok[:, :4]
    assert np.allclose(dev[:, 21:25][zok], host[:, 21:25][zok], rtol=1e-12, atol=1e-13)
    # the concentration c = g / exp(x) is evaluated in double on the device and in long double by the script / host reduction:
    # a bin holding a handful of pairs amplifies that 1e-16 by c^2 / var(c); 1e-12 from ten pairs on, 1e-9 below
    many = ok[:, 4] & (host[:, -1] >= 10)
    few = ok[:, 4] & (host[:, -1] < 10)
    assert np.allclose(dev[many, 25], host[many, 25], rtol=1e-12, atol=1e-13) and np.allclose(dev[few, 25], host[few, 25], rtol=1e-9, atol=1e-12)
    same = np.isclose(dev[:, 11:21:2], host[:, 11:21:2], rtol=0, atol=1e-15)
    assert same.mean() > 0.98
    assert np.allclose(dev[:, 12:21:2][same], host[:, 12:21:2][same], rtol=1e-9, atol=1e-300)
    assert np.allclose(dev[:, 1:11], host[:, 1:11], rtol=1e-6, atol=1e-12)


@pytest.mark.gpu
def test_correlation_sums_at_config4_size():
    """BASELINE configs[4] (100 k cells, two segments): the device reduction over all 52 M joints; pair counts of every lag
    equal the closed count (every point pairs once with each earlier point of its lineage at that lag)"""
    import gfp_gaussian_process_b200 as ggp
    P = ggp.PARAMS_SCALED_BINOMIAL
    P2 = np.stack([P, P * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    d = ggp.simulate_forest(1587, 6, params=P, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
    f = ggp.Forest(d)
    ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
    dt = 15.0
    sums, n_joints = ggp.api.correlation_sums(f, P2, dt, 200)
    assert n_joints == ggp.count_joints(f, P2, 1e-10)
    # lag k: one pair per point whose lineage reaches back k steps (uniform time grid: depth in points of every point)
    depth = np.zeros(d.n_ctp, dtype=np.int64)
    first = d.cell_offset[:-1]
    n = np.diff(d.cell_offset)
    for c in range(d.n_cells):   # parents precede daughters in the generator's order
        base = 0 if d.parent[c] < 0 else depth[d.cell_offset[d.parent[c] + 1] - 1] + 1
        depth[first[c]:first[c] + n[c]] = base + np.arange(n[c])
    want = np.array([d.n_ctp] + [(depth >= k).sum() for k in range(1, 200)])
    assert np.array_equal(np.asarray(sums[:, 0], dtype=np.int64), want)
    assert np.isfinite(np.asarray(sums, dtype=np.float64)).all()
    f.close()


@pytest.mark.gpu
def test_correlation_sums_take_any_number_of_lag_bins_and_repeat(monkeypatch):
    """the accumulators live in global memory (lock-free compensated sums): more lag bins than the 256 the first version's
    shared-memory bins held; the sums of the leading bins do not depend on the number of bins; two runs agree although
    the order of the atomic additions differs (the error words make the totals exact to second order)"""
    import gfp_gaussian_process_b200 as ggp
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(12, 5, params=P, noise_model="scaled", division_model="binomial", seed=5, pts_range=(6, 11))
    f = ggp.Forest(d)
    ggp.prediction_forward_backward(f, [P], forward=False, backward=False, combined=False)
    dt = float(d.time[1] - d.time[0])
    s40, n40 = ggp.api.correlation_sums(f, [P], dt, 40)
    s600, n600 = ggp.api.correlation_sums(f, [P], dt, 600)
    s600b, _ = ggp.api.correlation_sums(f, [P], dt, 600)
    assert n40 == n600 == ggp.count_joints(f, [P], 1e-10) > 0
    assert np.array_equal(s600[:, 0], s600b[:, 0]) and np.array_equal(s40[:, 0], s600[:40, 0])
    assert s600[:, 0].sum() > s40[:, 0].sum() or s600[40:, 0].sum() == 0      # deeper lags exist only beyond bin 40
    a, b, c = (np.asarray(x, dtype=np.float64) for x in (s40, s600[:40], s600b[:40]))
    assert np.allclose(a, b, rtol=1e-14, atol=1e-300) and np.allclose(b, c, rtol=1e-14, atol=1e-300)
    assert np.all(s600[int(s600[:, 0].nonzero()[0].max()) + 1:] == 0)
    # a record budget of 1 MB (2 600 records): many row blocks, and blocks that overflow their buffers are halved and walked again
    monkeypatch.setenv("GGP_B200_CORR_BUDGET", str(1 << 20))
    s_small, n_small = ggp.api.correlation_sums(f, [P], dt, 600)
    monkeypatch.delenv("GGP_B200_CORR_BUDGET")
    assert n_small == n600 and np.array_equal(s_small[:, 0], s600[:, 0])
    assert np.allclose(np.asarray(s_small, dtype=np.float64), np.asarray(s600, dtype=np.float64), rtol=1e-14, atol=1e-300)
    f.close()


@pytest.mark.gpu
def test_joints_into_preallocated_arrays():
    """collect_joint_distributions(out=...) fills the caller's (e.g. pinned) arrays with the same records"""
    import gfp_gaussian_process_b200 as ggp
    P = ggp.PARAMS_CONST_GAUSS
    d = ggp.simulate_forest(5, 4, seed=3, pts_range=(5, 8))
    f = ggp.Forest(d)
    ggp.prediction_forward_backward(f, [P], forward=False, backward=False, combined=False)
    r, c, m, v = ggp.collect_joint_distributions(f, [P], 1e-10)
    n = len(r)
    out = (np.full(n + 7, -1, dtype=np.int64), np.full(n + 7, -1, dtype=np.int64), np.full((n + 7, 44), np.nan))
    r2, c2, m2, v2 = ggp.collect_joint_distributions(f, [P], 1e-10, out=out)
    assert len(r2) == n and np.array_equal(r, r2) and np.array_equal(c, c2)
    assert np.array_equal(m.view(np.uint64), m2.view(np.uint64)) and np.array_equal(v.view(np.uint64), v2.view(np.uint64))
    assert np.all(out[0][n:] == -1)
    with pytest.raises(ValueError):
        ggp.collect_joint_distributions(f, [P], 1e-10, out=(out[0], out[1], np.zeros((n, 43))))
    f.close()
