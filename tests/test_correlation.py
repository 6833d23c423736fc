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
