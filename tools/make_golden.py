#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ from the REFERENCE ITSELF.  Run in the build
container, where /root/reference is mounted; the GPU box only sees the committed files.

  dawson_known_answers.json  the real-argument entries of Faddeeva.cc's own Dawson self-test (Maple values,
                             Faddeeva.cc:2378-2512) + values from mpmath at 60 decimal digits on a grid that
                             crosses every branch of w_im (Taylor, 97 Chebyshev pieces, continued fraction, 1/x)
  ref_step_vectors.npz       inputs and outputs of the reference's mean_cov_model / cross_cov_model / *tauint
                             (oracle/_ref = mean_cov_model.h + Faddeeva.cc compiled unmodified): the literal
                             inputs of tests.h:112-117, :138-166, :190-221 and states sampled along a filter run
                             on the example data set and on synthetic forests
  libm_bits.npz              exp/log/pow of the glibc the reference links (bit patterns) on the argument ranges the model visits
  example_forest.npz         example_data_set/input.csv parsed by the reference's rules (moma_input.h:401-527) and
                             the oracle's results on it (log-likelihood fresh / carried, per-cell sums, predictions
                             at a sample of points)
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from oracle import oracle_py  # noqa: E402
from gfp_gaussian_process_b200 import io as gio  # noqa: E402
from gfp_gaussian_process_b200.synthetic import simulate_forest, PARAMS_CONST_GAUSS, PARAMS_SCALED_BINOMIAL  # noqa: E402


def dawson_golden():
    src = open(os.path.join(REF, "src", "Faddeeva.cc")).read().split("\n")
    block = "\n".join(src[2377:2512])
    num = r"[-+]?(?:Inf|NaN|[0-9.]+(?:e[-+]?[0-9]+)?)"
    arrays = re.findall(r"cmplx [zw]\[NTST\] = \{(.*?)\};", block, flags=re.S)
    assert len(arrays) == 2
    vals = []
    for a in arrays:
        vals.append(re.findall(r"C\(\s*(" + num + r")\s*,\s*(" + num + r")\s*\)", a))
    z, w = vals
    assert len(z) == 48 and len(w) == 48, (len(z), len(w))
    maple = []
    for (zr, zi), (wr, wi) in zip(z, w):
        if float(zi.replace("Inf", "inf").replace("NaN", "nan")) == 0.0 and "N" not in zr and "I" not in zr:
            maple.append({"x": zr, "dawson": wr})
    import mpmath
    mpmath.mp.dps = 60
    grid = np.concatenate([np.linspace(-12, 12, 97), [1e-9, 1e-4, 0.0308, 0.031, 0.5, 0.92413, 1.65, 3.0, 7.7, 20.0, 44.9, 45.1,
                                                        1e3, 4.9e7, 5.1e7, 1e12, -1.65, -12.0, -46.0, -6e7]])
    mp = []
    for x in grid:
        xm = mpmath.mpf(float(x))
        d = mpmath.sqrt(mpmath.pi) / 2 * mpmath.exp(-xm * xm) * mpmath.erfi(xm)
        mp.append({"x": float(x).hex(), "dawson": mpmath.nstr(d, 30)})
    json.dump({"maple_real": maple, "mpmath": mp}, open(os.path.join(OUT, "dawson_known_answers.json"), "w"), indent=1)
    print("dawson: %d Maple (real-argument) + %d mpmath values" % (len(maple), len(mp)))


def sample_states(rng):
    """(mean[4], cov[16], dt, p7) tuples: tests.h literals + states along real filter runs"""
    cases = []
    # tests.h:190-221
    cov = np.zeros((4, 4))
    cov[0, 0], cov[1, 1], cov[2, 2], cov[3, 3] = 1, 2, 3, 4
    cov[1, 0] = cov[0, 1] = 2
    cov[3, 1] = 3
    cov[1, 3] = 4
    cases.append((np.array([1., 2, 3, 4]), cov.reshape(-1), 1.0, np.array([1., 2, 3, 4, 5, 6, 7])))
    # tests.h:138-166 (moment-term literals)
    up = [0.6, 0.7, 0.8, 0.9, 1, 1.1, 1.2, 1.3, 1.4, 1.5]
    c = np.zeros((4, 4))
    c[np.triu_indices(4)] = up
    c = c + np.triu(c, 1).T
    cases.append((np.array([0.2, 0.3, 0.4, 0.5]), c.reshape(-1), 0.1, np.array([1.6, 1.7, 1.8, 1.9, 2, 2.1, 2.2])))
    # tests.h:264-298 (realistic literals)
    c = np.array([4.25476409e-02, 4.81488709e+01, -6.17116203e-05, -1.25892662e-01, 4.81488709e+01, 1.67680116e+06,
                  2.59861605e-01, 7.45274531e+02, -6.17116203e-05, 2.59861605e-01, 8.48575294e-07, 8.44383560e-05,
                  -1.25892662e-01, 7.45274531e+02, 8.44383560e-05, 1.63738212e+00])
    cases.append((np.array([6.93147181e-01, 6.03801845e+03, 1.00811380e-02, 9.56031050e+00]), c, 15.0,
                  np.array([0.01, 0.01, 1e-05, 10, 0.01, 0.1, 0.001])))
    # states along filter runs: forward posteriors of the oracle
    cfg = gio.read_csv_config(os.path.join(REF, "example_data_set", "csv_config.txt"))
    ps = gio.read_parameter_file(os.path.join(REF, "example_data_set", "parameter_file.txt"))
    data, _ = gio.read_data(os.path.join(REF, "example_data_set", "input.csv"), cfg)
    P = np.array([p.init for p in ps])
    runs = [(data, P), (simulate_forest(6, 4, seed=3), PARAMS_CONST_GAUSS),
            (simulate_forest(6, 4, noise_model="scaled", division_model="binomial", seed=4), PARAMS_SCALED_BINOMIAL)]
    for d, p in runs:
        o = oracle_py.Oracle(d)
        pr = o.predictions([p])
        mf, cf = pr["forward"]
        mb, cb = pr["backward"]
        idx = rng.choice(d.n_ctp - 1, size=60, replace=False)
        for i in idx:
            dt = abs(d.time[i + 1] - d.time[i]) or 1.0
            cases.append((mf[i], cf[i].reshape(-1), dt, p[:7]))
        # backward frame: sign-flipped parameters and states (mean_cov_model_r, predictions.h:191-198)
        for i in idx[:20]:
            m = mb[i] * np.array([1, 1, -1, -1])
            c = cb[i] * np.array([[1, 1, -1, -1], [1, 1, -1, -1], [-1, -1, 1, 1], [-1, -1, 1, 1]])
            cases.append((m, c.reshape(-1), 1.0 if d is data else 3.5, p[:7] * np.array([-1, 1, 1, -1, 1, 1, -1])))
    return cases


def step_vectors():
    rng = np.random.default_rng(1)
    R = oracle_py.ref()
    assert R is not None, "reference build missing"
    cases = sample_states(rng)
    n = len(cases)
    mean = np.array([c[0] for c in cases])
    cov = np.array([c[1] for c in cases])
    dt = np.array([c[2] for c in cases])
    p7 = np.array([c[3] for c in cases])
    mo, co, cr = np.zeros((n, 4)), np.zeros((n, 16)), np.zeros((n, 16))
    for i in range(n):
        mo[i], co[i] = oracle_py.mean_cov_model(mean[i], cov[i], dt[i], p7[i], which="ref")
        cr[i] = oracle_py.cross_cov_model(mean[i], cov[i], dt[i], p7[i], which="ref")
    # integrals: tests.h:112-117 literal + arguments in the range the filter produces
    targs = [(0.0111, 0.022, 0.01, 0.2, 0.7)]
    for _ in range(200):
        a = 10 ** rng.uniform(-8, -1)
        b = rng.uniform(-0.05, 0.05)
        c = rng.uniform(-3, 3)
        t1 = rng.uniform(0.5, 15)
        targs.append((a, b, c, 0.0, t1))
        targs.append((a, b, c, t1, 2 * t1))
    targs = np.array(targs)
    tout = np.array([[R.ggp_ref_tauint(k, *row) for k in range(4)] for row in targs])
    np.savez_compressed(os.path.join(OUT, "ref_step_vectors.npz"), mean=mean, cov=cov, dt=dt, p7=p7, mean_out=mo, cov_out=co,
                        cross_out=cr, tauint_args=targs, tauint_out=tout)
    print("step vectors: %d states, %d integral argument tuples (non-finite outputs: %d)" %
          (n, len(targs), int((~np.isfinite(co)).any(axis=1).sum())))


def libm_bits():
    rng = np.random.default_rng(2)
    R = oracle_py.ref()
    xe = np.concatenate([rng.uniform(-40, 40, 4000), rng.uniform(-1e-3, 1e-3, 500), rng.uniform(-745, 709, 1500),
                         [0.0, -0.0, 1e-300, -1e-300, 709.9, -745.2, 800, -800, np.inf, -np.inf]])
    xl = np.concatenate([10 ** rng.uniform(-12, 12, 4000), rng.uniform(0.9, 1.1, 1500), [1.0, 5e-324, 1e-310, np.inf]])
    xp = np.concatenate([10 ** rng.uniform(-12, 2, 6000), rng.uniform(0.5, 2, 1000)])
    yp = rng.choice([1.5, 2.5, 3.5, 3.0], size=xp.shape[0])
    ex = np.array([R.ggp_ref_exp(v) for v in xe])
    lg = np.array([R.ggp_ref_log(v) for v in xl])
    pw = np.array([R.ggp_ref_pow(a, b) for a, b in zip(xp, yp)])
    np.savez_compressed(os.path.join(OUT, "libm_bits.npz"), exp_x=xe, exp_y=ex, log_x=xl, log_y=lg, pow_x=xp, pow_e=yp, pow_y=pw)
    print("libm: %d exp, %d log, %d pow" % (len(xe), len(xl), len(xp)))


def example_forest():
    cfg = gio.read_csv_config(os.path.join(REF, "example_data_set", "csv_config.txt"))
    ps = gio.read_parameter_file(os.path.join(REF, "example_data_set", "parameter_file.txt"))
    data, ids = gio.read_data(os.path.join(REF, "example_data_set", "input.csv"), cfg)
    P = np.array([p.init for p in ps])
    o = oracle_py.Oracle(data)
    ll1, cell_ll = o.total_loglik(P, per_cell=True)
    ll2 = o.total_loglik(P, fresh=False)
    ll3 = o.total_loglik(P, fresh=False)
    pr = o.predictions([P])
    sample = np.arange(0, data.n_ctp, 97)
    np.savez_compressed(
        os.path.join(OUT, "example_forest.npz"), cell_offset=data.cell_offset, parent=data.parent, time=data.time,
        log_length=data.log_length, fp=data.fp, params=P, loglik_fresh=ll1, loglik_second=ll2, loglik_third=ll3,
        cell_ll=cell_ll, sample=sample,
        **{f"{k}_{w}": pr[k][j][sample] for k in pr for j, w in enumerate(("mean", "cov"))})
    print("example: %d cells, %d ctp, loglik %r / %r / %r" % (data.n_cells, data.n_ctp, ll1, ll2, ll3))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    oracle_py.build()
    dawson_golden()
    step_vectors()
    libm_bits()
    example_forest()
