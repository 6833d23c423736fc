#!/usr/bin/env python
"""The reference's own +-1-ulp envelope: how far its results move when every exp, true pow and Dawson
RESULT inside mean_cov_model.h is nudged one ulp up or down at random.  That is the noise floor any implementation
that is not bit-identical is judged against (the fast likelihood kernel's gate is reported beside it; tolerances are never
widened to it).

The perturbed build is the reference's unmodified mean_cov_model.h + Faddeeva.cc with the three calls renamed by macros
(oracle/ref_shim.cpp, -DGGP_REF_ULP_PERTURB -> oracle/_ref/libggp_oracle_ulp.so) inside the oracle's filter loop.
CPU only; needs /root/reference (or the prebuilt library).

  python tools/ulp_envelope.py [--seeds 4] [--out profiles/r02_ulp_envelope.json]

Data sets: the example data set (scaled / binomial, parameter_file.txt values), a configs[1] sample (const / gauss,
1 500 trees x 6 generations), a scaled / binomial forest (400 trees x 6 generations).  Per data set and seed: relative
change of the total log-likelihood, of the per-cell sums (max), and of the forward / backward / combined predictions
(max over points, per mean and covariance entry).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gfp_gaussian_process_b200 as ggp  # noqa: E402
from oracle import oracle_py  # noqa: E402

NAMES = ["m_x", "m_g", "m_l", "m_q", "C_xx", "C_xg", "C_xl", "C_xq", "C_gg", "C_gl", "C_gq", "C_ll", "C_lq", "C_qq"]
IU = [(i, j) for i in range(4) for j in range(i, 4)]


def entries(mean, cov):
    return np.concatenate([mean, np.stack([cov[:, i, j] for i, j in IU], axis=1)], axis=1)


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def envelope(data, P, seeds, L, predictions=True):
    o = oracle_py.Oracle(data, lib=L)
    L.ggp_ref_set_ulp_seed(0)
    ll0, pc0 = o.total_loglik(P, per_cell=True)
    pr0 = o.predictions([P]) if predictions else None
    out = {"n_ctp": int(data.n_ctp), "n_cells": int(data.n_cells), "loglik": ll0, "seeds": []}
    for s in range(1, seeds + 1):
        L.ggp_ref_set_ulp_seed(1000003 * s)
        ll, pc = o.total_loglik(P, per_cell=True)
        row = {"seed": s, "loglik_rel": abs(ll - ll0) / abs(ll0), "cell_ll_rel_max": float(rel(pc, pc0).max())}
        if predictions:
            L.ggp_ref_set_ulp_seed(1000003 * s)
            pr = o.predictions([P])
            for k in ("forward", "backward", "prediction"):
                r = rel(entries(*pr[k]), entries(*pr0[k])).max(axis=0)
                row[k] = {n: float(v) for n, v in zip(NAMES, r)}
        out["seeds"].append(row)
    L.ggp_ref_set_ulp_seed(0)
    out["loglik_rel_max"] = max(r["loglik_rel"] for r in out["seeds"])
    if predictions:
        for k in ("forward", "backward", "prediction"):
            out[k + "_rel_max"] = {n: max(r[k][n] for r in out["seeds"]) for n in NAMES}
    return out


def datasets(small=False):
    from conftest import example_data
    ex, z = example_data(os.path.join(ROOT, "tests", "golden"))
    yield "example data set (scaled/binomial)", ex, np.asarray(z["params"])
    yield ("configs[1] sample (const/gauss)", ggp.simulate_forest(60 if small else 1500, 6, seed=20261018), ggp.PARAMS_CONST_GAUSS)
    yield ("scaled/binomial forest", ggp.simulate_forest(30 if small else 400, 6, noise_model="scaled", division_model="binomial",
                                                         seed=20261018), ggp.PARAMS_SCALED_BINOMIAL)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=4)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_ulp_envelope.json"))
    a = ap.parse_args()
    L = oracle_py.oracle_ulp()
    if L is None:
        raise SystemExit("oracle/_ref/libggp_oracle_ulp.so is not built and /root/reference is not mounted")
    res = {"what": __doc__.split("\n\n")[0], "seeds": a.seeds, "datasets": {}}
    for name, d, P in datasets(a.small):
        e = envelope(d, P, a.seeds, L)
        res["datasets"][name] = e
        print(f"{name}: {e['n_ctp']} ctp, loglik {e['loglik']!r}")
        print("   loglik rel change per seed:", " ".join(f"{r['loglik_rel']:.2e}" for r in e["seeds"]), " per-cell max:",
              " ".join(f"{r['cell_ll_rel_max']:.1e}" for r in e["seeds"]))
        for k in ("forward", "backward", "prediction"):
            print(f"   {k:10s}", " ".join(f"{n}={e[k + '_rel_max'][n]:.1e}" for n in NAMES))
    json.dump(res, open(a.out, "w"), indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
