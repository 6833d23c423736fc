"""BASELINE.json's own configurations at (or near) full size, and the code paths only they reach: vector-chunked batches
(cfg4: 4096 vectors x 630 000 cells does not fit one end-of-cell state buffer), carry chains across a chunk boundary,
the joints walk at cfg5 size with wrapped walker stacks.  GPU against the oracle, bit for bit where the quantity has no
reduction, north-star tolerances otherwise."""
import numpy as np
import pytest

from conftest import same_bits, max_rel
from oracle.oracle_py import Oracle
import gfp_gaussian_process_b200 as ggp

pytestmark = pytest.mark.gpu
LL_RTOL = 1e-10


def scan_vectors(P, n_vec):
    """cfg4's batch: 1-d sweeps of each parameter +-20 % around the truth"""
    per = -(-n_vec // 11)
    vecs = np.tile(P, (n_vec, 1))
    for k in range(n_vec):
        vecs[k, k // per] *= 0.8 + 0.4 * (k % per) / max(per - 1, 1)
    return vecs


@pytest.mark.parametrize("carry", [False, True])
def test_vector_chunked_batches_small_forest(carry, monkeypatch):
    """GGP_B200_STATE_BUDGET forces the 37-vector batch into chunks of 5 (8 chunks): fresh mode, and carry mode where the
    roots' chain has to cross every chunk boundary (main.cpp:102-108, predictions.h:64-78); per-cell sums and the returned
    root_carry against the oracle's sequential loop, bit for bit"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(40, 4, noise_model="scaled", division_model="binomial", seed=9)
    monkeypatch.setenv("GGP_B200_STATE_BUDGET", str(d.n_cells * 14 * 8 * 5))
    f = ggp.Forest(d)
    monkeypatch.delenv("GGP_B200_STATE_BUDGET")
    f1 = ggp.Forest(d)   # one chunk
    o = Oracle(d)
    rng = np.random.default_rng(1)
    vecs = P * (1 + 0.02 * rng.standard_normal((37, 11)))
    c = np.zeros((f.n_roots, 16)) if carry else None
    c1 = np.zeros((f.n_roots, 16)) if carry else None
    ll, pc = ggp.total_likelihood(vecs, f, root_carry=c, per_cell=True)
    ll1, pc1 = ggp.total_likelihood(vecs, f1, root_carry=c1, per_cell=True)
    assert same_bits(pc, pc1) and same_bits(ll, ll1)
    o.reset()
    for i in range(37):
        assert same_bits(pc[i], o.total_loglik(vecs[i], fresh=not carry, per_cell=True)[1]), i
    if carry:
        assert same_bits(c, c1) and same_bits(c, o.cell_cov[d.roots()])
    f.close()
    f1.close()


def test_cfg4_scan_4096_vectors_in_one_call():
    """BASELINE configs[3]: 4096 parameter vectors x the 10 000-tree forest in ONE ggp_loglik call (the library splits it
    into chunks of ~113 vectors to keep the state buffer under 8 GB).  All results finite; three spot vectors equal their
    single-vector calls bit for bit; every 128th vector, evaluated again as a chunked 32-vector batch with per-cell output,
    equals the big batch (totals) and the oracle on a 40-tree subset (per-cell sums), bit for bit"""
    P = ggp.PARAMS_CONST_GAUSS
    d = ggp.simulate_forest(10000, 6, seed=20261018)
    f = ggp.Forest(d)
    vecs = scan_vectors(P, 4096)
    ll = ggp.total_likelihood(vecs, f, raise_on_nan=False)
    assert ll.shape == (4096,) and np.isfinite(ll).all()
    for i in (0, 2049, 4095):
        assert ggp.total_likelihood(vecs[i], f) == ll[i]
    spot = np.arange(0, 4096, 128)
    ll_s, pc_s = ggp.total_likelihood(vecs[spot], f, per_cell=True)
    assert same_bits(ll_s, ll[spot])
    sub, cells, ctp = d.subset(d.roots()[4321:4361])
    o = Oracle(sub)
    for k, i in enumerate(spot):
        assert same_bits(pc_s[k][cells], o.total_loglik(vecs[i], per_cell=True)[1]), i
    f.close()


def test_cfg4_carry_chain_across_vector_chunks(monkeypatch):
    """carry mode at cfg4's forest size: 24 vectors in chunks of 7 (state budget forced), the roots' chain crossing three
    chunk boundaries; root_carry and per-cell sums of a 10-tree subset against the oracle's sequential loop"""
    P = ggp.PARAMS_CONST_GAUSS
    d = ggp.simulate_forest(10000, 6, seed=20261018)
    monkeypatch.setenv("GGP_B200_STATE_BUDGET", str(d.n_cells * 14 * 8 * 7))
    f = ggp.Forest(d)
    vecs = scan_vectors(P, 24)
    carry = np.zeros((f.n_roots, 16))
    ll, pc = ggp.total_likelihood(vecs, f, root_carry=carry, per_cell=True)
    roots = d.roots()
    sub, cells, ctp = d.subset(roots[7000:7010])
    o = Oracle(sub)
    o.reset()
    for i in range(24):
        assert same_bits(pc[i][cells], o.total_loglik(vecs[i], fresh=False, per_cell=True)[1]), i
    assert same_bits(carry[7000:7010], o.cell_cov[sub.roots()])
    f.close()


def test_cfg5_joints_at_full_size(monkeypatch):
    """BASELINE configs[4]: 1 587 trees x 6 generations (100 k cells, ~2 M ctp), two parameter segments, tol 1e-10.
    The count-only call over every start point equals the sum of the counts of row blocks; the records of a 6-tree block
    (rows, columns, 44 doubles) equal the oracle's on that subset bit for bit; the number of walker blocks (one per SM by
    default, here also 5 in total and two per SM) does not change a bit: the start-point counter, the per-walker pending
    stacks and the device sort are exercised where they wrap"""
    P = ggp.PARAMS_SCALED_BINOMIAL
    P2 = np.stack([P, P * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    d = ggp.simulate_forest(1587, 6, params=P, noise_model="scaled", division_model="binomial", seed=20261018, n_segments=2)
    assert d.n_cells == 99981
    f = ggp.Forest(d)
    ggp.prediction_forward_backward(f, P2, forward=False, backward=False, combined=False)
    total = ggp.api.count_joints(f, P2, 1e-10)
    assert total > 20 * d.n_ctp
    # counts of row blocks add up to the total
    edges = np.linspace(0, d.n_ctp, 8).astype(np.int64)
    assert sum(ggp.api.count_joints(f, P2, 1e-10, int(a), int(b)) for a, b in zip(edges[:-1], edges[1:])) == total
    # a 6-tree block against the oracle
    roots = d.roots()
    sub, cells, ctp = d.subset(roots[800:806])
    assert np.array_equal(ctp, np.arange(ctp[0], ctp[-1] + 1))   # trees are contiguous in the time-point order
    o = Oracle(sub)
    o.predictions(P2)
    n, row, col, rec = o.joints(1e-10, 4000000)
    assert n <= 4000000
    order = np.lexsort((col, row))
    r, c, mean, cov = ggp.collect_joint_distributions(f, P2, 1e-10, row_begin=int(ctp[0]), row_end=int(ctp[-1]) + 1)
    assert len(r) == n and np.array_equal(r - ctp[0], row[order]) and np.array_equal(c - ctp[0], col[order])
    assert same_bits(mean, rec[order][:, :8]) and same_bits(cov, rec[order][:, 8:])
    f.close()
    for env, val in (("GGP_B200_WALK_TOTAL_BLOCKS", "5"), ("GGP_B200_WALK_BLOCKS", "2")):
        monkeypatch.setenv(env, val)
        f2 = ggp.Forest(d)
        monkeypatch.delenv(env)
        ggp.prediction_forward_backward(f2, P2, forward=False, backward=False, combined=False)
        assert ggp.api.count_joints(f2, P2, 1e-10) == total
        r2, c2, m2, v2 = ggp.collect_joint_distributions(f2, P2, 1e-10, row_begin=int(ctp[0]), row_end=int(ctp[-1]) + 1)
        assert np.array_equal(r2, r) and np.array_equal(c2, c) and same_bits(m2, mean) and same_bits(v2, cov)
        f2.close()
