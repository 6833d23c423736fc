"""The FAST likelihood step (csrc/ggp_fast.cuh) compiled for the host (tests/hostcheck/fastcheck.cpp): its gate against the
oracle and its distance from a binary128 evaluation, without a GPU.  The device kernel runs the same code with FMA contraction."""
import os
import sys

import numpy as np
import pytest

from conftest import example_data, ROOT
from oracle.oracle_py import Oracle
import gfp_gaussian_process_b200 as ggp

sys.path.insert(0, os.path.join(ROOT, "tools"))
from fast_gate import fast_loglik  # noqa: E402

GATE = 1e-10


@pytest.mark.parametrize("noise,division,nodes", [("const", "gauss", 5), ("scaled", "binomial", 6), ("scaled", "gauss", 6), ("const", "binomial", 5)])
def test_fast_step_meets_the_gate_and_is_exact_against_binary128(noise, division, nodes):
    P = ggp.PARAMS_CONST_GAUSS if noise == "const" else ggp.PARAMS_SCALED_BINOMIAL
    d = ggp.simulate_forest(12, 4, params=P, noise_model=noise, division_model=division, seed=71)
    ref, pc_ref = Oracle(d).total_loglik(P, per_cell=True)
    fast, valid, pc, _ = fast_loglik(d, P, n_nodes=nodes, per_cell=True)
    quad, _, pc_q, _ = fast_loglik(d, P, quad=True, per_cell=True)
    assert valid[0] == 1
    assert abs(fast[0] - ref) / abs(ref) <= GATE
    assert abs(fast[0] - quad[0]) / abs(quad[0]) <= 1e-14 and np.max(np.abs(pc[0] - pc_q[0]) / np.abs(pc_q[0])) <= 1e-12
    # the reference itself is further from the binary128 value than the fast step is
    assert abs(ref - quad[0]) >= abs(fast[0] - quad[0])


def test_fast_step_on_the_example_data_set(golden_dir):
    data, z = example_data(golden_dir)
    P = np.asarray(z["params"])
    ref = Oracle(data).total_loglik(P)
    fast, valid, _, _ = fast_loglik(data, P, n_nodes=5)
    assert valid[0] == 1 and abs(fast[0] - ref) / abs(ref) <= GATE


def test_fast_step_flags_steps_outside_its_validity_range():
    P = ggp.PARAMS_CONST_GAUSS.copy()
    d = ggp.simulate_forest(3, 3, seed=72)
    assert fast_loglik(d, P, n_nodes=5)[1][0] == 1
    P[4] = 2.0   # gamma_q = 2: the exponent varies by 7 over a step
    assert fast_loglik(d, P, n_nodes=5)[1][0] == 0 and fast_loglik(d, P, n_nodes=10)[1][0] == 0
    P[4] = 0.25
    assert fast_loglik(d, P, n_nodes=6)[1][0] == 0 and fast_loglik(d, P, n_nodes=8)[1][0] == 1
