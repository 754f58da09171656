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
    return sorted(set(re.findall(r'\b(omg_[a-z_]+)\s*\(', src)))


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
    assert lib.omg_abi_version() == 4


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
