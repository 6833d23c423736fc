"""The drop-in, end to end (GPU): the reference's own unmodified headers build std::vector<MOMAdata> (genealogy, init_cells),
the reference-side binding include/ggp_bridge.h flattens it and calls libggp_b200.so, the results go back into the
reference's own containers and through the reference's own writers - compared with the reference's CPU passes on the very
same objects (oracle/_ref/libggp_ref_bridge.so, built from /root/reference in the build container; it travels to the GPU box)."""
import numpy as np
import pytest

from conftest import same_bits, ragged_forest
from oracle.oracle_py import ref_wrappers, RefWrappers
import gfp_gaussian_process_b200 as ggp

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref_wrappers(bridge=True) is None, reason="oracle/_ref/libggp_ref_bridge.so not built")]
LL_RTOL = 1e-10   # north star: log-likelihood within relative 1e-10 (the total is a different summation order; per-cell sums are bit-equal elsewhere)


@pytest.mark.parametrize("noise,division", [("const", "gauss"), ("scaled", "binomial")])
def test_binding_matches_the_reference_cpu_passes(noise, division):
    P = ggp.PARAMS_CONST_GAUSS if noise == "const" else ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(6, 4, params=P, noise_model=noise, division_model=division, seed=41, pts_range=(4, 8))
    r = RefWrappers(d, bridge=True)
    r.open_bridge()
    vecs = np.stack([P, P * 1.01, P, P * 0.98])
    # four successive evaluations: the reference's history-dependent chain (H3) on the CPU ...
    r.reset()
    cpu = np.array([r.total_loglik(v, fresh=False) for v in vecs])
    # ... through the binding's nlopt-signature objective, and as one batched call on a second bridge (fresh carry)
    gpu = r.bridge_loglik(vecs)
    assert np.max(np.abs(gpu - cpu) / np.abs(cpu)) <= LL_RTOL
    r.close()
    r = RefWrappers(d, bridge=True)
    r.open_bridge()
    assert same_bits(r.bridge_loglik(vecs, batch=True), gpu)
    # predictions: reference CPU passes, then the binding filling the same per-cell vectors
    want = r.predictions([P])
    text_cpu = r.joints_text(1e-10, gpu=False, precision=17)
    got = r.bridge_predictions([P])
    for k in ("forward", "backward", "prediction"):
        assert same_bits(got[k][0], want[k][0]) and same_bits(got[k][1], want[k][1]), k
    # joints: the reference's dense CSV text, at 17 digits (bit-exact) and at the writer's default 6
    assert r.joints_text(1e-10, gpu=True, precision=17) == text_cpu
    assert r.joints_text(1e-10, gpu=True, precision=6) == r.joints_text(1e-10, gpu=False, precision=6)
    r.close()


def test_binding_on_segments_and_ragged_trees():
    d = ragged_forest()
    P = np.stack([ggp.PARAMS_SCALED_BINOMIAL, ggp.PARAMS_SCALED_BINOMIAL * np.array([1, 1, 1, 2, 1, 1, 1, 1, 1, 1, 1.])])
    r = RefWrappers(d, bridge=True)
    r.open_bridge()
    want = r.predictions(P)
    text_cpu = r.joints_text(1e-10, gpu=False, precision=17)
    got = r.bridge_predictions(P)
    for k in ("forward", "backward", "prediction"):
        assert same_bits(got[k][0], want[k][0]) and same_bits(got[k][1], want[k][1]), k
    assert r.joints_text(1e-10, gpu=True, precision=17) == text_cpu
    r.close()
