"""The CPU oracle (oracle/ipm_ref.py): pinned against committed golden vectors,
KKT conditions and an independent SLSQP optimum.  ("parity unpinned" w.r.t.
IPOPT itself: no CasADi/IPOPT binary exists in this image.)"""
import os

import numpy as np
import pytest

from omg_tools_b200 import scenarios as sc
from oracle import ipm_ref
from oracle.nlp_eval import TableEval

G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'p2p_golden.npz'))


@pytest.fixture(scope='module')
def cfg1():
    return sc.config1(build_solver=False)


def test_oracle_reproduces_golden_config1(cfg1):
    tb = cfg1.father.tables
    X0, P = sc.instance_data(cfg1, 4, jitter=0.2, seed=1)
    assert np.array_equal(X0, G['config1_X0']) and np.array_equal(P, G['config1_P'])
    for b in (0, 2):
        r = ipm_ref.solve(tb, X0[b], P[b])
        assert r.status == 0 and r.iters == G['config1_loose_iters'][b]
        assert np.abs(r.x - G['config1_loose_x'][b]).max() < 1e-9
        assert np.abs(r.lam_g - G['config1_loose_lam'][b]).max() < 1e-8


def test_oracle_kkt_conditions_at_tight_tolerance(cfg1):
    tb = cfg1.father.tables
    ev = TableEval(tb)
    x, lam = G['config1_tight_x'][0], G['config1_tight_lam'][0]
    V = ev.tape(G['config1_P'][0])
    g = ev.g(x, V)
    eq = tb.lbg == tb.ubg
    assert np.abs(g[eq] - tb.lbg[eq]).max() < 1e-7           # feasibility
    assert g[~eq].max() < 1e-7
    stat = ev.gradf(x, V) + ev.jac_dense(x, V).T @ lam
    assert np.abs(stat).max() < 1e-6                          # stationarity
    assert lam[~eq].min() > -1e-9                             # dual feasibility
    assert np.abs(lam[~eq] * g[~eq]).max() < 1e-6             # complementarity


def test_oracle_optimum_matches_independent_slsqp():
    # vehicle spline coefficients are unique; hyperplanes a,b are not
    assert abs(G['config1_tight_f'][0] - float(G['config1_slsqp_f'])) < 1e-6
    assert np.abs(G['config1_tight_x'][0][:26] - G['config1_slsqp_x'][:26]).max() < 1e-5
    # default tolerance (tol=1e-3) stops on the central path near mu ~ 1e-5
    assert np.abs(G['config1_loose_x'][0][:26] - G['config1_slsqp_x'][:26]).max() < 1e-3


def test_signed_cholesky_solves_permuted_saddle_system():
    """K = L S L^T with equality rows interleaved after their variables."""
    rng = np.random.default_rng(0)
    n, ne = 12, 3
    A = rng.standard_normal((n, n))
    H = A @ A.T + n * np.eye(n)
    Jc = np.zeros((ne, n))
    for k in range(ne):
        Jc[k, 3 * k:3 * k + 4] = rng.standard_normal(4)
    K = np.block([[H, Jc.T], [Jc, np.zeros((ne, ne))]])
    order = list(range(4)) + [n] + list(range(4, 7)) + [n + 1] + \
        list(range(7, 10)) + [n + 2] + list(range(10, n))
    sign = np.array([1.0 if v < n else -1.0 for v in order])
    Kp = K[np.ix_(order, order)]
    ok, L, eqf = ipm_ref._signed_cholesky(Kp, sign, 1e-12)
    assert ok and not eqf
    assert np.abs(L @ np.diag(sign) @ L.T - Kp).max() < 1e-10
    rhs = rng.standard_normal(n + ne)
    u = np.empty(n + ne)
    u[order] = ipm_ref._signed_solve(L, sign, rhs[order])
    assert np.abs(K @ u - rhs).max() < 1e-10
    Kbad = Kp.copy()
    Kbad[0, 0] -= 1000.0
    assert not ipm_ref._signed_cholesky(Kbad, sign, 1e-12)[0]


def test_kkt_structure_is_consistent(cfg1):
    tb = cfg1.father.tables
    N = tb.kkt_n
    pos = np.r_[tb.kkt_pos_var, tb.kkt_pos_eq]
    assert sorted(pos) == list(range(N))
    assert np.all(tb.env_first[:N] % 8 == 0) and np.all(tb.env_first[:N] <= np.arange(N))
    assert tb.env_size == tb.env_ptr[-1] and tb.env_size < (N + 1) * (N + 2) // 2
    # every equality row sits after all variables it couples (negative pivot)
    for k, i in enumerate(tb.kkt_eq_rows):
        cols = tb.jcol[tb.jrow_ptr[i]:tb.jrow_ptr[i + 1]]
        assert tb.kkt_pos_eq[k] > tb.kkt_pos_var[cols].max()
    # every H entry and border entry lies inside the envelope
    for q in range(tb.nnz_h):
        a, b = tb.kkt_pos_var[tb.hrow[q]], tb.kkt_pos_var[tb.hcol[q]]
        hi, lo = max(a, b), min(a, b)
        assert lo >= tb.env_first[hi]
        assert tb.kkt_hdst[q] == tb.env_ptr[hi] + lo - tb.env_first[hi]
