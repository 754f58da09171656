"""The C-ABI shared library loads and exports every function declared in
include/omg_b200.h (no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'omg_b200.h')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(omg_[a-z0-9_]+)\s*\(', src)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    from omg_tools_b200.solver import b200
    return b200.load_library()


def test_exports_match_header(lib):
    from omg_tools_b200.solver import b200
    names = declared_functions()
    assert set(names) == set(b200.EXPORTS)
    for name in names:
        assert hasattr(lib, name), name
    assert lib.omg_abi_version() == 6


def test_default_options_are_the_reference_ipopt_settings(lib):
    from omg_tools_b200.solver.b200 import _Options
    o = _Options()
    lib.omg_default_options(ctypes.byref(o))
    assert o.tol == 1e-3                # reference problem.py:57
    assert o.mu_init == 0.1 and o.max_iter == 3000
    assert o.constr_viol_tol == 1e-4 and o.compl_inf_tol == 1e-4


def test_no_cpu_fallback_without_device(lib):
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from omg_tools_b200 import scenarios as sc
    from omg_tools_b200.solver.b200 import B200Solver
    pr = sc.config1(build_solver=False)
    with pytest.raises(RuntimeError, match='no CUDA device|CUDA'):
        B200Solver(pr.father.tables)
    with pytest.raises(RuntimeError):
        sc.config1(build_solver=True)


def test_table_file_round_trip(tmp_path):
    """save_tables -> omg_tables_read (host only): the deployable artefact for
    native callers reproduces every array of the lowered NLP, including the
    intermediates of config 4 and the cross-Hessian lists of the default Dubins
    formulation."""
    import ctypes as C
    import numpy as np
    from omg_tools_b200 import scenarios as sc
    from omg_tools_b200.solver import b200
    lib = b200.load_library()
    for builder in (sc.config1, sc.config4, sc.config_dubins_plain):
        tb = builder(build_solver=False).father.tables
        path = str(tmp_path / 'problem.omgtbl')
        b200.save_tables(tb, path)
        T = lib.omg_tables_read(path.encode())
        assert bool(T), lib.omg_last_error()
        t = T.contents
        assert (t.abi_version, t.n, t.m, t.n_par) == (b200.ABI_VERSION, tb.n, tb.m, tb.n_par)
        assert (t.n_mid, t.nnz_j, t.kkt_n, t.env_size) == (
            getattr(tb, 'n_mid', 0), tb.nnz_j, tb.kkt_n, tb.env_size)
        arr = lambda ptr, cnt: np.ctypeslib.as_array(ptr, (cnt,))
        for name in ('G', 'J', 'W'):
            src, dst = getattr(tb, name), getattr(t, name)
            assert np.array_equal(arr(dst.coef, dst.n_terms), src.coef)
            assert np.array_equal(arr(dst.xi, dst.n_terms * dst.width), src.xi.reshape(-1))
            assert np.array_equal(arr(dst.ptr, dst.n_out + 1), src.ptr)
        assert np.array_equal(arr(t.W.lrow, t.W.n_terms), tb.W.lrow)
        assert np.array_equal(arr(t.lbg, t.m), tb.lbg)
        assert np.array_equal(arr(t.kkt_hdst, t.nnz_h), tb.kkt_hdst)
        if t.n_mid:
            assert np.array_equal(arr(t.jp_a, t.n_jp), tb.jp_a)
            assert np.array_equal(arr(t.mu_slot, t.n_mu), tb.mu_slot)
        assert t.nnz_wx == getattr(tb, 'nnz_wx', 0) and t.W.n_out == t.nnz_w + t.nnz_wx
        if t.nnz_wx:
            assert np.array_equal(arr(t.xq_h, t.n_xq), tb.xq_h)
            assert np.array_equal(arr(t.xq_ptr, t.n_xq + 1), tb.xq_ptr)
            assert np.array_equal(arr(t.xq_w, t.n_xp), tb.xq_w)
            assert np.array_equal(arr(t.xq_a, t.n_xp), tb.xq_a)
            assert np.array_equal(arr(t.xq_b, t.n_xp), tb.xq_b)
        lib.omg_tables_free(T)
    # a damaged file is refused with a message
    with open(path, 'r+b') as fp:
        fp.write(b'XXXX')
    assert not bool(lib.omg_tables_read(path.encode()))
    assert b'table file' in lib.omg_last_error()
