import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """compile the checkers (oracle, host check) and make sure the product library exists"""
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    """bit-for-bit equality; any NaN equals any NaN"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def example_data(golden_dir):
    from gfp_gaussian_process_b200 import LineageData
    z = np.load(os.path.join(golden_dir, "example_forest.npz"))
    data = LineageData(cell_offset=z["cell_offset"], parent=z["parent"], time=z["time"], log_length=z["log_length"], fp=z["fp"],
                       noise_model="scaled", division_model="binomial")
    return data, z


def ragged_forest():
    """two parameter segments, single-daughter mothers, one-point cells, parents stored after daughters, fp_auto != 0"""
    import gfp_gaussian_process_b200 as ggp
    rng = np.random.default_rng(3)
    d = ggp.simulate_forest(5, 4, noise_model="scaled", division_model="binomial", seed=13, n_segments=2, pts_range=(1, 6))
    keep = np.ones(d.n_cells, dtype=bool)
    for c in rng.choice(np.flatnonzero(d.daughter2 >= 0), size=6, replace=False):
        stack = [int(d.daughter2[c])]
        while stack:
            u = stack.pop()
            keep[u] = False
            stack += [int(k) for k in (d.daughter1[u], d.daughter2[u]) if k >= 0]
    cells = np.flatnonzero(keep)[::-1]
    remap = -np.ones(d.n_cells, dtype=np.int64)
    remap[cells] = np.arange(len(cells))
    n = np.diff(d.cell_offset)[cells]
    off = np.concatenate([[0], np.cumsum(n)])
    ctp = np.repeat(d.cell_offset[cells] - off[:-1], n) + np.arange(off[-1])
    par = d.parent[cells]
    return ggp.LineageData(cell_offset=off, parent=np.where(par >= 0, remap[np.maximum(par, 0)], -1), time=d.time[ctp],
                           log_length=d.log_length[ctp], fp=d.fp[ctp], segment=d.segment[ctp], noise_model="scaled",
                           division_model="binomial", fp_auto=3.0)
