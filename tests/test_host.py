"""Host logic and the C-ABI surface, CPU only: the shared library loads and exports everything include/ggp_b200.h
declares, fails loudly without a GPU, and the host-side mirrors of the reference's helpers behave like them."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import _lib, sharding, api
from gfp_gaussian_process_b200.forest import build_daughters


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ggp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ggp_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES), "ctypes binding and header disagree"
    assert b"sm_100a" in lib.ggp_version()


def test_library_is_sm100a_and_not_linked_to_oracle():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "ggp_ref" not in ldd
    # the product sources never mention the oracle
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gfp_gaussian_process_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".inc")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("oracle/", "ORACLE_DIR/") or True
    src = "".join(open(os.path.join(ROOT, "gfp_gaussian_process_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "gfp_gaussian_process_b200")) if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src and "oracle_py" not in src


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    d = ggp.simulate_forest(2, 2)
    with pytest.raises(_lib.GgpError) as e:
        ggp.Forest(d)
    assert e.value.code == _lib.GGP_ERR_CUDA


def test_bad_arguments_are_rejected():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ggp_forest_create(None, C.byref(h)) == _lib.GGP_ERR_BAD_ARG
    d = ggp.simulate_forest(2, 2)
    from hostpass import make_desc
    desc = make_desc(d)
    desc.n_cells = 0
    assert lib.ggp_forest_create(C.byref(desc), C.byref(h)) == _lib.GGP_ERR_BAD_ARG
    assert b"empty" in lib.ggp_last_error()
    assert lib.ggp_loglik(None, None, 1, None, None, None, None) == _lib.GGP_ERR_BAD_ARG
    # three children of one mother: build_cell_genealogy throws (moma_input.h:141-147)
    bad = ggp.LineageData(cell_offset=np.arange(5), parent=np.array([-1, 0, 0, 0]), time=np.arange(4.), log_length=np.zeros(4),
                          fp=np.ones(4), daughter1=np.array([1, -1, -1, -1]), daughter2=np.array([2, -1, -1, -1]))
    assert lib.ggp_forest_create(C.byref(make_desc(bad)), C.byref(h)) == _lib.GGP_ERR_BAD_ARG


def test_build_daughters_is_file_order():
    parent = np.array([-1, 0, 5, 0, -1, 4, 5])
    d1, d2 = build_daughters(parent)
    assert d1.tolist() == [1, -1, -1, -1, 5, 2, -1] and d2.tolist() == [3, -1, -1, -1, -1, 6, -1]


def test_arange_accumulates_like_the_reference():
    # utils.h:96-103 adds the step repeatedly; 0.1 * 3 != 0.1 + 0.1 + 0.1
    a = api.arange(0.0, 1.0, 0.1)
    v, ref = 0.0, []
    while v < 1.0:
        ref.append(v)
        v += 0.1
    assert a.tolist() == ref and len(a) == 11 and a[8] != 0.1 * 8 and len(np.arange(0.0, 1.0, 0.1)) == 10


def test_hessian_stencil_order():
    x = np.arange(1.0, 12.0)
    vecs, hs = api.hessian_stencil(x, [0, 7], 1e-3)
    assert vecs.shape == (16, 11) and len(hs) == 4
    assert np.allclose(vecs[0], x + np.eye(11)[0] * 2e-3)          # i = j = 0: +h +h on the same entry
    assert np.allclose(vecs[5], x + np.eye(11)[0] * 1e-3 - np.eye(11)[7] * 8e-3)


def test_init_stats_match_layout_code(golden_dir):
    from conftest import example_data
    from oracle.oracle_py import Oracle
    data, _ = example_data(golden_dir)
    f, r = data.init_stats()
    of, orr = Oracle(data).init_stats()
    assert np.array_equal(f, of) and np.array_equal(r, orr)


def test_partition_is_balanced_and_complete():
    d = ggp.simulate_forest(37, 3, seed=4)
    parts = sharding.partition_roots(d, 4)
    allr = np.sort(np.concatenate(parts))
    assert np.array_equal(allr, d.roots())
    roots, size = sharding.tree_sizes(d)
    loads = [size[np.isin(roots, p)].sum() for p in parts]
    assert max(loads) - min(loads) <= size.max()
    sub, cells, ctp = sharding.shard(d, 1, 4)
    assert sub.n_ctp == loads[1] and np.array_equal(sub.time, d.time[ctp])
    assert np.array_equal(sub.init_f, d.init_stats()[0])   # population statistics of the WHOLE data set


def test_sharded_loglik_sums_to_total_gloo_world2(tmp_path):
    """world_size-2 run over gloo: each rank evaluates its shard (host bodies of the product code here, the kernels
    on a GPU box), one all-reduce of the per-vector sums; equals the single-process total to rounding"""
    script = tmp_path / "w.py"
    script.write_text(f"""
import os, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import numpy as np, torch, torch.distributed as dist
import gfp_gaussian_process_b200 as ggp
from gfp_gaussian_process_b200 import sharding
from hostpass import host_loglik
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
d = ggp.simulate_forest(16, 3, seed=8)
P = np.stack([ggp.PARAMS_CONST_GAUSS, ggp.PARAMS_CONST_GAUSS * 1.01])
sub, cells, ctp = sharding.shard(d, rank, world)
cl, _ = host_loglik(sub, P)
local = torch.from_numpy(cl.sum(axis=1))
tot = sharding.allreduce_loglik(local.clone())
full, _ = host_loglik(d, P)
if rank == 0:
    ref = full.sum(axis=1)
    assert np.all(np.abs(tot.numpy() - ref) <= 1e-12 * np.abs(ref)), (tot, ref)
    # per-cell values are unchanged by sharding (the init statistics are frozen before the split)
assert np.array_equal(cl, full[:, cells])
dist.destroy_process_group()
print('rank', rank, 'ok')
""")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_io_readers(tmp_path):
    from gfp_gaussian_process_b200 import io as gio
    (tmp_path / "cfg.txt").write_text("# c\ntime_col = t\nlength_col = len\nfp_col = fl\nrescale_time = 60\ncell_tags = lane, id\nparent_tags = lane, pid\n")
    (tmp_path / "p.txt").write_text("mean_lambda = 1, 0.1\ngamma_lambda = 2\nvar_lambda = 3, 0.1, 1, 5\nmean_q = 4\ngamma_q = 5\nvar_q = 6\nbeta = 7\nvar_x = 8\nvar_g = 9\nvar_dx = 10\nvar_dg = 11\n")
    (tmp_path / "d.csv").write_text("t,len,fl,lane,id,pid\n0,2.0,10,1,1.0,0\n60,2.5,11,1,1.0,0\n120,1.2,5,1,2,1\n120,1.3,6,1,3,1.0\n180,1.4,7,1,3,1\n")
    cfg = gio.read_csv_config(str(tmp_path / "cfg.txt"))
    ps = gio.read_parameter_file(str(tmp_path / "p.txt"))
    assert [p.kind for p in ps[:3]] == ["free", "fixed", "bound"] and ps[2].upper == 5
    data, ids = gio.read_data(str(tmp_path / "d.csv"), cfg)
    assert ids == ["1.1", "1.2", "1.3"] and data.parent.tolist() == [-1, 0, 0]
    assert data.daughter1.tolist() == [1, -1, -1] and data.daughter2.tolist() == [2, -1, -1]
    assert data.time.tolist() == [0, 1, 2, 2, 3] and data.log_length[0] == np.log(2.0)


def test_upload_chunks_keep_trees_together_and_cover_the_series():
    """ggp_layout.hpp: the series are cut into contiguous chunks (equal, or of given relative sizes); a tree belongs to the
    chunk its LAST point arrives with, so every cell of a tree sits in one chunk and all its points have landed when that
    chunk has; the chunk boundaries are monotone and cover [0, n_ctp)."""
    import ctypes as C
    from hostpass import hc, make_desc
    data = ggp.simulate_forest(37, 4, seed=3)
    desc = make_desc(data)
    root = np.arange(data.n_cells)
    for c in range(data.n_cells):          # parents precede daughters in the generator's order
        if data.parent[c] >= 0:
            root[c] = root[data.parent[c]]
    last = data.cell_offset[1:] - 1
    for want, fr in ((1, None), (3, None), (7, None), (4, [1.0, 5.0, 2.0, 2.0]), (3, [1e-9, 1.0, 1.0])):
        starts = np.zeros(want + 1, dtype=np.int64)
        chunk = np.full(data.n_cells, -1, dtype=np.int32)
        frp = (C.c_double * want)(*fr) if fr else None
        k = hc().hc_layout_chunks(C.byref(desc), want, frp, starts.ctypes.data_as(C.POINTER(C.c_longlong)),
                                  chunk.ctypes.data_as(C.POINTER(C.c_int)))
        assert k == want
        assert starts[0] == 0 and starts[-1] == data.n_ctp and np.all(np.diff(starts) >= 0)
        if fr:
            w = np.cumsum([0.0] + fr) / sum(fr)
            assert np.all(np.abs(starts - w * data.n_ctp) <= 1)
        assert np.all(chunk >= 0)
        for r in np.unique(root):
            cells = np.nonzero(root == r)[0]
            assert len(set(chunk[cells])) == 1                               # a tree is never split
            k_tree = chunk[cells[0]]
            assert last[cells].max() < starts[k_tree + 1]                    # all its points have landed with that chunk
            assert last[cells].max() >= starts[k_tree]                       # and not earlier: it is the chunk of the last point
