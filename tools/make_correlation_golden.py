#!/usr/bin/env python
"""Golden vectors for the correlation post-processing (host/ggp_correlation.hpp) from the REFERENCE's own script
python_src/correlation_from_joint.py, imported here (it cannot travel to the GPU box).  Inputs: the prediction and
dense joints files of a small forest, written in the reference's file format from the CPU oracle's results
(tests/corr_files.py); outputs: tests/golden/correlation_reference.npz with the script's per-lag results.
usage: python tools/make_correlation_golden.py   (needs /root/reference)"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from corr_files import CASES, write_case  # noqa: E402

# the script imports matplotlib at module level (plot helpers we do not use); give it an empty stand-in
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
    if name not in sys.modules:
        m = types.ModuleType(name)
        sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
if not hasattr(np, "longfloat"):
    np.longfloat = np.longdouble   # alias removed in NumPy 2.0; the script predates it
spec = importlib.util.spec_from_file_location("correlation_from_joint", "/root/reference/python_src/correlation_from_joint.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
for name, case in CASES.items():
    tmp = tempfile.mkdtemp()
    jf, pf, dt = write_case(case, tmp)
    for norm in (False, True):
        if norm and case["pts"][0] < 2:
            continue   # the script divides by a zero cell-cycle time for one-point cells (ZeroDivisionError)
        if norm:
            import pandas as pd
            cells = ref.df2ggp_cells(pd.read_csv(pf, skiprows=ref.header_lines(pf)))
            corr = ref.files2correlation_function(jf, pf, np.arange(0, 3, 0.05), 0.024, normalize_time=True,
                                                  cell_cylce_time=ref.get_cell_cycle_times(cells))
        else:
            corr = ref.files2correlation_function(jf, pf, np.arange(0, dt * case["n_data"], dt), dt * 0.2)
        csv = os.path.join(tmp, "ref.csv")
        ref.corr_to_csv(corr, csv)
        rows = [l.rstrip(",\n").split(",") for l in open(csv).read().strip().split("\n")[1:]]
        key = name + ("_norm" if norm else "")
        out[key + "_table"] = np.array([[float(x) for x in r] for r in rows])
        out[key + "_n"] = np.array([c.n for c in corr])
        out[key + "_cov"] = np.array([c.cov for c in corr])
        out[key + "_cov_c"] = np.array([c.cov_concentration for c in corr])
        print(key, "lags", len(corr), "pairs", out[key + "_n"][:6], "...")
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "correlation_reference.npz"), **out)
print("written")
