"""Parity of the CUDA path, called through the C ABI, against the oracle and the committed golden vectors.
Bars (BASELINE.json north_star): relative 1e-10 on the log-likelihood, 1e-9 on predicted means/covariances.
What is actually asserted is stronger wherever the arithmetic is order-fixed: bit-for-bit equality of per-cell
log-likelihood sums, of every stored mean/covariance, of exp/log/pow/Dawson and of the propagation step."""
import json
import os

import numpy as np
import pytest

from conftest import same_bits, max_rel, example_data, ragged_forest
from oracle.oracle_py import Oracle
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import api

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-10      # north_star bar for the log-likelihood
PRED_RTOL = 1e-9     # north_star bar for predicted means / covariances


def test_device_math_bits(golden_dir):
    z = np.load(os.path.join(golden_dir, "libm_bits.npz"))
    assert same_bits(api.math_eval("exp", z["exp_x"]), z["exp_y"])
    assert same_bits(api.math_eval("log", z["log_x"]), z["log_y"])
    assert same_bits(api.math_eval("pow", z["pow_x"], z["pow_e"]), z["pow_y"])
    g = json.load(open(os.path.join(golden_dir, "dawson_known_answers.json")))
    xs = np.array([float(e["x"]) for e in g["maple_real"]])
    ws = np.array([float(e["dawson"]) for e in g["maple_real"]])
    assert max_rel(api.math_eval("dawson", xs), ws) < 1e-13
    xs = np.array([float.fromhex(e["x"]) for e in g["mpmath"]])
    L = __import__("oracle.oracle_py", fromlist=["oracle"]).oracle()
    assert same_bits(api.math_eval("dawson", xs), np.array([L.ggp_oracle_dawson(v) for v in xs]))
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-60, 60, 50000), 10 ** rng.uniform(-12, 9, 5000)])
    assert same_bits(api.math_eval("dawson", xs), np.array([L.ggp_oracle_dawson(v) for v in xs]))


def test_fast_paths_of_the_cooperative_step_fall_back_exactly():
    """The cooperative step evaluates pow with a role's exponentials as one interleaved block and finishes the
    log-evidence term inline (shared-reciprocal division, log's main path); every argument outside those main paths
    must go through the out-of-line routines and give the same bits: ordinary, tiny, huge, zero, negative, inf, nan."""
    import ctypes as C
    from hostpass import hc
    rng = np.random.default_rng(11)
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7e308,
                   1e-20, 1e-17, 600.0, -600.0, 720.0, -745.2, 1100.0, -1100.0, 1 - 2.0 ** -53, 1 + 2.0 ** -52, 0.97, 1.03])
    # pow + exp block
    base = np.concatenate([10 ** rng.uniform(-12, 4, 20000), rng.uniform(0.9, 1.1, 2000), sp, rng.permutation(sp)])
    arg = np.concatenate([rng.uniform(-40, 40, 20000), 10 ** rng.uniform(-25, 3, 2000) * rng.choice([-1, 1], 2000), rng.permutation(sp), sp])
    got = api.math_eval("powexp", base, arg)
    e = 1.5 + (np.arange(base.size) % 3)
    with np.errstate(all="ignore"):
        ref = np.stack([api.math_eval("pow", base, e), api.math_eval("exp", arg), api.math_eval("exp", arg / 2),
                        api.math_eval("exp", -arg), api.math_eval("exp", arg + 1)], axis=1)
    assert same_bits(got, ref)
    # log-evidence term: random well-conditioned S, S with a row exchange, det near 1, singular / negative / non-finite S
    n = 20000
    S00 = 10 ** rng.uniform(-4, 6, n); S11 = 10 ** rng.uniform(-4, 6, n)
    rho = rng.uniform(-0.99, 0.99, n)
    S01 = rho * np.sqrt(S00 * S11); S10 = S01 * (1 + rng.choice([0, 2.0 ** -52, -2.0 ** -52], n))
    qf = -rng.uniform(0, 50, n)
    cases = np.stack([qf, S00, S01, S10, S11], axis=1)
    near1 = np.stack([qf[:500], 1 + rng.uniform(-0.05, 0.05, 500), np.zeros(500), np.zeros(500), np.ones(500)], axis=1)
    a3, b3 = [v.reshape(-1) for v in np.meshgrid(sp, sp)]
    odd = np.stack([np.full(a3.size, -1.5), a3, b3, b3[::-1], a3[::-1]], axis=1)
    cases = np.ascontiguousarray(np.concatenate([cases, near1, odd]))
    ref = np.empty(cases.shape[0])
    hc().hc_ll_finish(C.c_long(cases.shape[0]), cases.ctypes.data_as(C.POINTER(C.c_double)), ref.ctypes.data_as(C.POINTER(C.c_double)))
    assert same_bits(api.math_eval("ll_finish", cases), ref)


def test_shared_reciprocal_division_is_ieee():
    """a / GgpDivisor(b) must be the correctly rounded quotient (what the host's `/` returns) for every operand
    class: the step's ~190 divisions per time point go through it"""
    rng = np.random.default_rng(4)
    n = 400000
    a = rng.standard_normal(n) * 10 ** rng.uniform(-300, 300, n)
    b = rng.standard_normal(n) * 10 ** rng.uniform(-300, 300, n)
    a2 = rng.standard_normal(n) * 10 ** rng.uniform(-12, 12, n)
    b2 = rng.standard_normal(n) * 10 ** rng.uniform(-12, 12, n)
    sp = np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1.7e308, 2.2250738585072014e-308, 3.0, 1 / 3])
    a3, b3 = [v.reshape(-1) for v in np.meshgrid(sp, sp)]
    a = np.concatenate([a, a2, a3, b2 * (1 + 2.0 ** -52)])
    b = np.concatenate([b, b2, b3, b2])
    with np.errstate(all="ignore"):
        ref = a / b
    assert same_bits(api.math_eval("div", a, b), ref)


def test_device_step_matches_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "ref_step_vectors.npz"))
    iu = np.triu_indices(4)
    s14 = np.concatenate([z["mean"], z["cov"].reshape(-1, 4, 4)[:, iu[0], iu[1]]], axis=1)
    out, cross = api.propagate_eval(s14, z["dt"], z["p7"], cross=True)
    assert same_bits(out[:, :4], z["mean_out"])
    assert same_bits(out[:, 4:], z["cov_out"].reshape(-1, 4, 4)[:, iu[0], iu[1]])
    assert same_bits(cross, z["cross_out"])


@pytest.mark.parametrize("noise,division", [("const", "gauss"), ("scaled", "binomial"), ("scaled", "gauss"), ("const", "binomial")])
def test_loglik_and_predictions_match_oracle(noise, division):
    P = ggp.PARAMS_CONST_GAUSS if noise == "const" else ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(300, 5, params=P, noise_model=noise, division_model=division, seed=5)
    f = ggp.Forest(d)
    o = Oracle(d)
    ll_ref, pc_ref = o.total_loglik(P, per_cell=True)
    ll, pc = ggp.total_likelihood(P, f, per_cell=True)
    assert same_bits(pc, pc_ref)                       # every cell's own sum, bit for bit
    assert abs(ll - ll_ref) <= LL_RTOL * abs(ll_ref)   # the total differs only by summation order
    assert abs(ll - ll_ref) <= 1e-13 * abs(ll_ref)
    pr, ref = ggp.prediction_forward_backward(f, [P]), o.predictions([P])
    for k in ("forward", "backward", "prediction"):
        for a, b in zip(pr[k], ref[k]):
            assert max_rel(a, b) <= PRED_RTOL
            assert same_bits(a, b)
    bm, bc = api.backward_cell_state(f)
    assert same_bits(bm, o.cell_mean) and same_bits(bc.reshape(-1, 16), o.cell_cov)
    f.close()


def test_example_dataset_golden(golden_dir):
    """the reference's own example data (config 1): 1 tree, 77 cells, depth 12, 22 065 points"""
    data, z = example_data(golden_dir)
    f = ggp.Forest(data)
    ll, pc = ggp.total_likelihood(z["params"], f, per_cell=True)
    assert same_bits(pc, z["cell_ll"])
    assert abs(ll - float(z["loglik_fresh"])) <= 1e-13 * abs(ll)
    carry = np.zeros((1, 16))
    l1 = ggp.total_likelihood(z["params"], f, root_carry=carry)
    l2 = ggp.total_likelihood(z["params"], f, root_carry=carry)
    l3 = ggp.total_likelihood(z["params"], f, root_carry=carry)
    for got, key in ((l1, "loglik_fresh"), (l2, "loglik_second"), (l3, "loglik_third")):
        assert abs(got - float(z[key])) <= 1e-13 * abs(got)
    pr = ggp.prediction_forward_backward(f, [z["params"]])
    for k in ("forward", "backward", "prediction"):
        assert same_bits(pr[k][0][z["sample"]], z[f"{k}_mean"])
        assert same_bits(pr[k][1][z["sample"]], z[f"{k}_cov"])
    f.close()


def test_carry_chain_and_batching():
    """a batch evaluates each vector exactly as a single call would; with root_carry the batch reproduces the
    reference's sequential history dependence"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(40, 4, noise_model="scaled", division_model="binomial", seed=9)
    f = ggp.Forest(d)
    o = Oracle(d)
    rng = np.random.default_rng(1)
    vecs = P * (1 + 0.02 * rng.standard_normal((37, 11)))
    ll_b, pc_b = ggp.total_likelihood(vecs, f, per_cell=True)
    for i in (0, 5, 36):
        ll_1, pc_1 = ggp.total_likelihood(vecs[i], f, per_cell=True)
        assert same_bits(pc_1, pc_b[i]) and ll_1 == ll_b[i]
        assert same_bits(pc_1, o.total_loglik(vecs[i], per_cell=True)[1])
    carry = np.zeros((f.n_roots, 16))
    ll_c, pc_c = ggp.total_likelihood(vecs[:6], f, root_carry=carry, per_cell=True)
    o.reset()
    for i in range(6):
        assert same_bits(pc_c[i], o.total_loglik(vecs[i], fresh=False, per_cell=True)[1])
    assert same_bits(carry, o.cell_cov[d.roots()])
    # chaining across two calls equals one chained batch
    carry2 = np.zeros((f.n_roots, 16))
    a = ggp.total_likelihood(vecs[:3], f, root_carry=carry2)
    b = ggp.total_likelihood(vecs[3:6], f, root_carry=carry2)
    assert np.array_equal(np.concatenate([a, b]), ll_c) and same_bits(carry2, carry)
    f.close()


def test_packed_prediction_outputs_are_the_writer_columns():
    """ggp_predict14: 4 means + upper triangle in the order write_predictions_to_file prints them (predictions.h:575-578),
    packed on the device; same bits as the corresponding entries of the full outputs"""
    d = ragged_forest()
    P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    f = ggp.Forest(d)
    full = ggp.prediction_forward_backward(f, P)
    packed = ggp.prediction_upper14(f, P)
    iu = [(i, j) for i in range(4) for j in range(i, 4)]
    for k in ("forward", "backward", "prediction"):
        assert same_bits(packed[k][:, :4], full[k][0])
        for c, (i, j) in enumerate(iu):
            assert same_bits(packed[k][:, 4 + c], full[k][1][:, i, j]), (k, i, j)
    only = ggp.prediction_upper14(f, P, forward=False, backward=False)
    assert list(only) == ["prediction"] and same_bits(only["prediction"], packed["prediction"])
    f.close()


def test_segments_ragged_and_single_point_cells():
    d = ragged_forest()
    P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    f = ggp.Forest(d)
    o = Oracle(d)
    pr, ref = ggp.prediction_forward_backward(f, P), o.predictions(P)
    for k in ("forward", "backward", "prediction"):
        assert same_bits(pr[k][0], ref[k][0]) and same_bits(pr[k][1], ref[k][1])
    assert same_bits(ggp.total_likelihood(P[0], f, per_cell=True)[1], o.total_loglik(P[0], per_cell=True)[1])
    f.close()


def test_nan_is_reported_like_the_reference():
    d = ggp.simulate_forest(30, 3, seed=2)
    P = ggp.PARAMS_CONST_GAUSS.copy()
    P[7] = -1.0
    f = ggp.Forest(d)
    o = Oracle(d)
    assert np.isnan(o.total_loglik(P))
    with pytest.raises(ggp.LikelihoodNaN) as e:
        ggp.total_likelihood(P, f)
    assert (e.value.cell, e.value.t_index) == o.nan
    good = ggp.total_likelihood(np.stack([ggp.PARAMS_CONST_GAUSS, P]), f, raise_on_nan=False)
    assert np.isfinite(good[0]) and np.isnan(good[1])
    f.close()


def test_scan_and_hessian_batches_match_oracle():
    P = ggp.PARAMS_CONST_GAUSS
    d = ggp.simulate_forest(20, 3, seed=6)
    f = ggp.Forest(d)
    o = Oracle(d)
    samp, ll = ggp.run_bound_1dscan(f, P, 3, 8.0, 12.0, 0.25)
    assert len(samp) == 16
    for s, l in zip(samp, ll):
        q = P.copy()
        q[3] = s
        ref = o.total_loglik(q)
        assert abs(l - ref) <= LL_RTOL * abs(ref)
    H = ggp.num_hessian_ll(f, P, [0, 3, 7], 1e-3)
    vecs, hs = api.hessian_stencil(P, [0, 3, 7], 1e-3)
    ref = np.array([o.total_loglik(v) for v in vecs]).reshape(-1, 4)
    Href = np.array([(r[0] - r[1] - r[2] + r[3]) / (4 * h1 * h2) for r, (h1, h2) in zip(ref, hs)]).reshape(3, 3)
    assert np.allclose(H, Href, rtol=1e-6, atol=1e-6 * np.abs(Href).max())
    f.close()


def test_full_size_properties():
    """BASELINE.json config 2 at full size (10 000 trees x 6 generations, ~12.6 M ctp), checked through size-
    independent properties: run-to-run determinism, additivity over a split of the trees, and bit-exact per-cell
    sums against the oracle on a 40-tree sample"""
    from gfp_gaussian_process_b200 import sharding
    P = ggp.PARAMS_CONST_GAUSS
    d = ggp.simulate_forest(10000, 6, seed=20261018)
    assert d.n_cells == 630000 and 12.0e6 < d.n_ctp < 13.2e6
    f = ggp.Forest(d)
    ll, pc = ggp.total_likelihood(P, f, per_cell=True)
    ll2 = ggp.total_likelihood(P, f)
    assert np.isfinite(ll) and ll == ll2
    assert abs(pc.sum() - ll) <= 1e-12 * abs(ll)
    halves = []
    for r in range(2):
        sub, cells, ctp = sharding.shard(d, r, 2)
        fs = ggp.Forest(sub)
        l_s, pc_s = ggp.total_likelihood(P, fs, per_cell=True)
        assert same_bits(pc_s, pc[cells])
        halves.append(l_s)
        fs.close()
    assert abs(sum(halves) - ll) <= 1e-12 * abs(ll)
    sub, cells, ctp = d.subset(d.roots()[1234:1274])
    o = Oracle(sub)
    ll_o, pc_o = o.total_loglik(P, per_cell=True)
    assert same_bits(pc[cells], pc_o)
    f.close()


@pytest.mark.parametrize("noise,division", [("const", "gauss"), ("scaled", "binomial")])
def test_joints_match_oracle(noise, division):
    """-j through the C ABI: every emitted joint P(z_col, z_row | D), 8 means + 36 covariances, against the oracle"""
    P = ggp.PARAMS_CONST_GAUSS if noise == "const" else ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(6, 4, params=P, noise_model=noise, division_model=division, seed=15, pts_range=(4, 8))
    f = ggp.Forest(d)
    o = Oracle(d)
    ggp.prediction_forward_backward(f, [P])
    o.predictions([P])
    n, row, col, rec = o.joints(1e-10, 400000)
    order = np.lexsort((col, row))
    r, c, mean, cov = ggp.collect_joint_distributions(f, [P], 1e-10)
    assert len(r) == n and np.array_equal(r, row[order]) and np.array_equal(c, col[order])
    assert max_rel(mean, rec[order][:, :8]) <= PRED_RTOL
    assert same_bits(mean, rec[order][:, :8]) and same_bits(cov, rec[order][:, 8:])
    # a row block gives exactly the rows asked for
    k0, k1 = d.n_ctp // 3, 2 * d.n_ctp // 3
    r2, c2, m2, v2 = ggp.collect_joint_distributions(f, [P], 1e-10, row_begin=k0, row_end=k1)
    sel = (r >= k0) & (r < k1)
    assert np.array_equal(r2, r[sel]) and np.array_equal(c2, c[sel]) and same_bits(m2, mean[sel])
    f.close()


def test_joints_segments_ragged():
    d = ragged_forest()
    P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    f = ggp.Forest(d)
    o = Oracle(d)
    ggp.prediction_forward_backward(f, P)
    o.predictions(P)
    n, row, col, rec = o.joints(1e-10, 400000)
    order = np.lexsort((col, row))
    r, c, mean, cov = ggp.collect_joint_distributions(f, P, 1e-10)
    assert len(r) == n and np.array_equal(r, row[order]) and np.array_equal(c, col[order])
    assert same_bits(mean, rec[order][:, :8]) and same_bits(cov, rec[order][:, 8:])
    with pytest.raises(Exception):
        ggp.collect_joint_distributions(f, P * 1.01, 1e-10)   # not the parameters of the prediction
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [1, 3, 7, "1,5,2,2"])
def test_streamed_upload_evaluates_chunk_by_chunk(chunks, monkeypatch):
    """ggp_forest_upload_series cuts the series into chunks; the next evaluation runs a chunk's trees as soon as the chunk
    has landed.  Same per-cell sums as a forest created from the new series directly, whatever the chunking; also with
    parents stored after daughters and one-point cells (ragged forest)."""
    if isinstance(chunks, str):   # chunks of unequal size (relative sizes), each on its own stream like the equal ones
        monkeypatch.setenv("GGP_B200_CHUNK_FRACTIONS", chunks)
    else:
        monkeypatch.setenv("GGP_B200_UPLOAD_CHUNKS", str(chunks))
    for d, P in ((ggp.simulate_forest(40, 4, seed=31), ggp.PARAMS_CONST_GAUSS), (ragged_forest(), ggp.PARAMS_SCALED_BINOMIAL)):
        f = ggp.Forest(d)
        ll0, pc0 = ggp.total_likelihood(P, f, per_cell=True)
        # new measurements on the same genealogy
        rng = np.random.default_rng(5)
        x2 = d.log_length + rng.normal(0, 1e-3, d.n_ctp)
        g2 = d.fp * (1 + rng.normal(0, 1e-3, d.n_ctp))
        t2 = d.time.copy()
        f.upload_series(t2.ctypes.data, x2.ctypes.data, g2.ctypes.data)
        vecs = np.stack([P, P * 1.01])
        ll1, pc1 = ggp.total_likelihood(vecs, f, per_cell=True)
        # the root / leaf priors follow the NEW measurements (init_cells_f/r, moma_input.h:675-735): the oracle derives its own
        d2 = ggp.LineageData(cell_offset=d.cell_offset, parent=d.parent, time=t2, log_length=x2, fp=g2, segment=d.segment,
                             noise_model=d.noise_model, division_model=d.division_model, fp_auto=d.fp_auto)
        o = Oracle(d2)
        assert same_bits(f.init_stats()[0], o.init_stats()[0]) and same_bits(f.init_stats()[1], o.init_stats()[1])
        for i in range(2):
            assert same_bits(pc1[i], o.total_loglik(vecs[i], per_cell=True)[1])
        assert not same_bits(pc1[0], pc0)
        # and again resident (no upload pending): same numbers
        ll2, pc2 = ggp.total_likelihood(vecs, f, per_cell=True)
        # (per-cell sums are identical; the forest total is reduced over different block partials, last-bit differences)
        assert same_bits(pc2, pc1) and max_rel(ll2, ll1) < 1e-14
        # predictions after an upload wait for all chunks
        # (only the measurements travel: the time grid is unchanged)
        f.upload_series(None, d.log_length.ctypes.data, d.fp.ctypes.data)
        pr = ggp.prediction_forward_backward(f, [P] * (int(d.segment.max()) + 1))
        ref = Oracle(d).predictions([P] * (int(d.segment.max()) + 1))
        assert same_bits(pr["prediction"][0], ref["prediction"][0])
        f.close()
        # a shard carries the statistics of the whole data set: new measurements must come with new statistics
        sub, cells, ctp = d.subset(d.roots()[:2])
        fs = ggp.Forest(sub)
        with pytest.raises(Exception):
            fs.upload_series(None, x2[ctp].copy().ctypes.data, None)
        o2 = Oracle(d2)
        xs, gs = np.ascontiguousarray(x2[ctp]), np.ascontiguousarray(g2[ctp])
        fs.upload_series(None, xs.ctypes.data, gs.ctypes.data, *o2.init_stats())
        sub2, _, _ = d2.subset(d2.roots()[:2])
        assert same_bits(ggp.total_likelihood(P, fs, per_cell=True)[1], Oracle(sub2).total_loglik(P, per_cell=True)[1])
        fs.close()


@pytest.mark.gpu
def test_sharded_predictions_and_joints_gather_to_the_unsharded_result():
    """trees shard across ranks (bin packing by cell-timepoints, descendants stay with their root, the
    init_cells statistics are those of the whole data set); -p / -j outputs stay sharded by ctp and gather by index.
    Three shards evaluated on this GPU reproduce the single-forest result bit for bit."""
    from gfp_gaussian_process_b200 import sharding
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(11, 4, noise_model="scaled", division_model="binomial", seed=41, pts_range=(3, 9))
    full = ggp.Forest(d)
    ll_full, pc_full = ggp.total_likelihood(P, full, per_cell=True)
    pr_full = ggp.prediction_forward_backward(full, [P])
    r_full, c_full, m_full, v_full = ggp.collect_joint_distributions(full, [P], 1e-10)
    world = 3
    parts = sharding.partition_roots(d, world)
    assert sorted(np.concatenate(parts).tolist()) == d.roots().tolist()
    pc = np.full(d.n_cells, np.nan)
    pred = {k: (np.full((d.n_ctp, 4), np.nan), np.full((d.n_ctp, 4, 4), np.nan)) for k in ("forward", "backward", "prediction")}
    joints = []
    for rank in range(world):
        sub, cells, ctp = sharding.shard(d, rank, world)
        f = ggp.Forest(sub)
        _, pcs = ggp.total_likelihood(P, f, per_cell=True)
        pc[cells] = pcs
        pr = ggp.prediction_forward_backward(f, [P])
        for k in pred:
            pred[k][0][ctp] = pr[k][0]
            pred[k][1][ctp] = pr[k][1]
        r, c, m, v = ggp.collect_joint_distributions(f, [P], 1e-10)
        joints.append((ctp[r], ctp[c], m, v))
        f.close()
    assert same_bits(pc, pc_full)
    for k in pred:
        assert same_bits(pred[k][0], pr_full[k][0]) and same_bits(pred[k][1], pr_full[k][1])
    r = np.concatenate([j[0] for j in joints]); c = np.concatenate([j[1] for j in joints])
    m = np.concatenate([j[2] for j in joints]); v = np.concatenate([j[3] for j in joints])
    order = np.lexsort((c, r))
    assert np.array_equal(r[order], r_full) and np.array_equal(c[order], c_full)
    assert same_bits(m[order], m_full) and same_bits(v[order], v_full)
    full.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ragged", [False, True])
def test_group_handle_shards_one_data_set_over_devices(ragged):
    """ggp_group: one handle, several shards (here three on the visible devices, wrapping around): per-cell sums, the carry
    chain, NaN reports and prediction rows come back in the caller's order, bit-identical to the single-forest results;
    tree-major data is split into contiguous runs (rows copied straight into place), other data is bin-packed and scattered"""
    import torch
    nd = max(torch.cuda.device_count(), 1)
    devices = [k % nd for k in range(3)]
    if ragged:
        d = ragged_forest()
        P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    else:
        d = ggp.simulate_forest(23, 4, noise_model="scaled", division_model="binomial", seed=77)
        P = np.stack([ggp.PARAMS_SCALED_BINOMIAL])
    g = ggp.ForestGroup(d, devices)
    assert g.size == 3 and g.contiguous == (not ragged)
    assert np.array_equal(np.sort(np.concatenate([g.member_ctp(k) for k in range(3)])), np.arange(d.n_ctp))
    f = ggp.Forest(d)
    vecs = np.stack([P[0], P[0] * 1.02, P[0] * 0.97])
    ll_g, pc_g = g.total_likelihood(vecs, per_cell=True)
    ll_f, pc_f = ggp.total_likelihood(vecs, f, per_cell=True)
    assert same_bits(pc_g, pc_f) and max_rel(ll_g, ll_f) < 1e-14
    c_g, c_f = np.zeros((f.n_roots, 16)), np.zeros((f.n_roots, 16))
    a = g.total_likelihood(vecs, root_carry=c_g, per_cell=True)[1]
    b = ggp.total_likelihood(vecs, f, root_carry=c_f, per_cell=True)[1]
    assert same_bits(a, b) and same_bits(c_g, c_f)
    bad = P[0].copy()
    bad[7] = -1.0
    g.total_likelihood(np.stack([P[0], bad]))
    o = Oracle(d)
    assert np.isnan(o.total_loglik(bad)) and g.nan[0] == (-1, -1) and g.nan[1] == o.nan
    full = ggp.prediction_forward_backward(f, P)
    got = g.predictions(P)
    got14 = g.predictions(P, packed=True)
    iu = [4 * i + j for i in range(4) for j in range(i, 4)]
    for k in ("forward", "backward", "prediction"):
        assert same_bits(got[k][:, :4], full[k][0]) and same_bits(got[k][:, 4:], full[k][1].reshape(-1, 16))
        assert same_bits(got14[k][:, :4], full[k][0]) and same_bits(got14[k][:, 4:], full[k][1].reshape(-1, 16)[:, iu])
    # joints and the lag-binned correlation sums over the shards: every shard walks its own trees; rows / columns are the
    # caller's time-point indices in (row, col) order, the sums add over shards
    r, c, m, v = ggp.collect_joint_distributions(f, P, 1e-10)
    rg, cg, mg, vg = g.joints(P, 1e-10)
    assert len(r) > 100 and np.array_equal(r, rg) and np.array_equal(c, cg) and same_bits(m, mg) and same_bits(v, vg)
    lo, hi = d.n_ctp // 3, 2 * d.n_ctp // 3
    sel = (r >= lo) & (r < hi)
    rb, cb, mb, vb = g.joints(P, 1e-10, row_begin=lo, row_end=hi)
    assert np.array_equal(rb, r[sel]) and np.array_equal(cb, c[sel]) and same_bits(vb, v[sel])
    if not ragged:   # (the ragged toy stores a parent after its daughter: the device reduction refuses it, the host path remains)
        dt = float(d.time[1] - d.time[0])
        s_f, n_f = ggp.api.correlation_sums(f, P, dt, 60)
        s_g, n_g = g.correlation_sums(P, dt, 60)
        assert n_f == n_g == len(r) and np.array_equal(s_f[:, 0], s_g[:, 0])
        assert np.allclose(np.asarray(s_f, dtype=np.float64), np.asarray(s_g, dtype=np.float64), rtol=1e-14, atol=1e-300)
    g.set_mode("fast")
    assert max_rel(g.total_likelihood(vecs), ll_f) <= 1e-10
    g.close()
    f.close()


@pytest.mark.gpu
@pytest.mark.parametrize("ng4_min", [None, "1"])
def test_cooperative_kernels_repeat_bit_for_bit_under_load(ng4_min, monkeypatch):
    """Evidence in place of compute-sanitizer's racecheck, which is closed on this pool: the cooperative kernels (named barriers,
    overlaid scratch regions, cp.async double buffer) give the SAME BITS in every repetition - per-cell sums, forward / backward /
    combined predictions, joints - while a second thread keeps the GPU busy with other launches of the same kernels (different
    co-residency and timing in every repetition); once with the default launch selection and once with the four-groups-per-block
    kernels forced (GGP_B200_NG4_MIN=1).  The first repetition equals the oracle (a subset of the trees; all cells in the other
    tests); a race in the scratch protocol would have to be invisible in 30 repetitions x 58 000 cell steps each."""
    import threading
    if ng4_min:
        monkeypatch.setenv("GGP_B200_NG4_MIN", ng4_min)
    P = ggp.PARAMS_SCALED_BINOMIAL
    P2 = np.stack([P, P * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    d = ggp.simulate_forest(300, 4, params=P, noise_model="scaled", division_model="binomial", seed=91, n_segments=2, pts_range=(9, 17))
    load = ggp.simulate_forest(2000, 5, seed=92)
    f, g = ggp.Forest(d), ggp.Forest(load)
    vecs = np.stack([P, P * 1.01, P * 0.98])
    stop = threading.Event()

    def busy():
        while not stop.is_set():
            ggp.total_likelihood(ggp.PARAMS_CONST_GAUSS, g)
            ggp.prediction_forward_backward(g, [ggp.PARAMS_CONST_GAUSS], forward=False, backward=False, combined=False)

    t = threading.Thread(target=busy)
    t.start()
    try:
        ll0, pc0 = ggp.total_likelihood(vecs, f, per_cell=True)
        pr0 = ggp.prediction_forward_backward(f, P2)
        j0 = ggp.collect_joint_distributions(f, P2, 1e-8, row_begin=0, row_end=400)
        for rep in range(30):
            ll, pc = ggp.total_likelihood(vecs, f, per_cell=True)
            assert same_bits(pc, pc0) and same_bits(ll, ll0), rep
            if rep % 3 == 0:
                pr = ggp.prediction_forward_backward(f, P2)
                for k in ("forward", "backward", "prediction"):
                    assert same_bits(pr[k][0], pr0[k][0]) and same_bits(pr[k][1], pr0[k][1]), (rep, k)
            if rep % 10 == 0:
                j = ggp.collect_joint_distributions(f, P2, 1e-8, row_begin=0, row_end=400)
                assert all(np.array_equal(a, b) for a, b in zip(j[:2], j0[:2])) and same_bits(j[2], j0[2]) and same_bits(j[3], j0[3])
    finally:
        stop.set()
        t.join()
    sub, cells, ctp = d.subset(d.roots()[:12])
    sub.init_f, sub.init_r = d.init_stats()
    o = Oracle(sub)
    assert same_bits(o.total_loglik(P, per_cell=True)[1], pc0[0][cells])
    f.close()
    g.close()


@pytest.mark.gpu
def test_group_handles_come_and_go():
    """a group owns one worker thread per shard: created, used and destroyed repeatedly (also unused, and with more shards than
    trees: empty shards have no member and no work), results unchanged"""
    d = ggp.simulate_forest(3, 3, seed=5)
    f = ggp.Forest(d)
    want = ggp.total_likelihood(ggp.PARAMS_CONST_GAUSS, f, per_cell=True)[1]
    f.close()
    for rep in range(12):
        g = ggp.ForestGroup(d, [0] * (1 + rep % 5))
        if rep % 3:
            for _ in range(3):
                assert same_bits(g.total_likelihood(ggp.PARAMS_CONST_GAUSS, per_cell=True)[1], want)
        g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [101, 102, 103, 104])
def test_random_forests_all_passes_bitwise(seed):
    """random shapes (tree counts that do not fill a 32-cell group, very short and long cells, 1-3 segments, both models):
    per-cell sums, forward / backward / combined predictions and the backward cell state against the oracle, bit for bit"""
    rng = np.random.default_rng(seed)
    noise, division = [("const", "gauss"), ("scaled", "binomial"), ("scaled", "gauss"), ("const", "binomial")][seed % 4]
    P0 = ggp.PARAMS_CONST_GAUSS if noise == "const" else ggp.PARAMS_SCALED_BINOMIAL
    n_seg = int(rng.integers(1, 4))
    d = ggp.simulate_forest(int(rng.integers(1, 70)), int(rng.integers(1, 6)), params=P0, noise_model=noise, division_model=division,
                            seed=seed, pts_range=(1, int(rng.integers(2, 40))), n_segments=n_seg, fp_auto=float(rng.integers(0, 3)))
    P = np.stack([P0 * (1 + 0.05 * rng.standard_normal(11)) for _ in range(n_seg)])
    f = ggp.Forest(d)
    o = Oracle(d)
    if n_seg == 1:
        vecs = np.stack([P[0], P[0] * 1.02, P[0] * 0.97])
        ll, pc = ggp.total_likelihood(vecs, f, per_cell=True, raise_on_nan=False)
        for i in range(3):
            assert same_bits(pc[i], o.total_loglik(vecs[i], per_cell=True)[1])
    pr, ref = ggp.prediction_forward_backward(f, P), o.predictions(P)
    for k in ("forward", "backward", "prediction"):
        assert same_bits(pr[k][0], ref[k][0]) and same_bits(pr[k][1], ref[k][1]), k
    bm, bc = ggp.api.backward_cell_state(f)
    assert same_bits(bm, o.cell_mean) and same_bits(bc.reshape(-1, 16), o.cell_cov)
    f.close()


@pytest.mark.gpu
def test_full_size_predictions_properties():
    """BASELINE.json config 3 at scale (scaled noise + binomial division, -p): 4 000 trees x 6 generations (252 000 cells,
    ~5 M cell-timepoints, 2.4 GB of outputs) checked through size-independent properties: every prediction finite with
    positive variances, run-to-run bitwise determinism, and bit-exact agreement with the oracle on a 30-tree sample
    (the init statistics being those of the whole forest), for all three outputs"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(4000, 6, noise_model="scaled", division_model="binomial", seed=20261018)
    f = ggp.Forest(d)
    pr = ggp.prediction_forward_backward(f, [P])
    for k in ("forward", "backward", "prediction"):
        assert np.isfinite(pr[k][0]).all() and np.isfinite(pr[k][1]).all()
        assert (np.diagonal(pr[k][1], axis1=1, axis2=2) > 0).all()
    again = ggp.prediction_forward_backward(f, [P], forward=False, backward=False)
    assert same_bits(again["prediction"][0], pr["prediction"][0]) and same_bits(again["prediction"][1], pr["prediction"][1])
    # the smoothed variances never exceed the filtered ones by more than rounding (information only adds)
    vf = np.diagonal(pr["forward"][1], axis1=1, axis2=2)[:, :2]
    vs = np.diagonal(pr["prediction"][1], axis1=1, axis2=2)[:, :2]
    assert (vs <= vf * (1 + 1e-6)).mean() > 0.999
    sub, cells, ctp = d.subset(d.roots()[777:807])
    ref = Oracle(sub).predictions([P])
    for k in ("forward", "backward", "prediction"):
        assert same_bits(pr[k][0][ctp], ref[k][0]) and same_bits(pr[k][1][ctp], ref[k][1]), k
    f.close()
