"""ORACLE: ctypes wrapper of oracle/_build/libipm_oracle.so (C restatement of
oracle/ipm_ref.py).  Test infrastructure / CPU baseline only."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, '_build', 'libipm_oracle.so')
_lib = None


def available():
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        vp = C.c_void_p
        _lib.oracle_solve_batch.argtypes = [vp, vp, C.c_int] + [vp] * 4 + [C.c_int] + \
            [vp] * 6 + [C.c_int]
        _lib.oracle_feas_batch.argtypes = [vp, C.c_int] + [vp] * 4 + [C.c_int, C.c_int] + [vp] * 3
    return _lib


def solve_batch(tb, X0, P, threads=0, options=None, lbg=None, ubg=None, lam_g0=None):
    """Returns (X, status, iters); extra results via return_all=True variant."""
    out = solve_batch_full(tb, X0, P, threads, options, lbg, ubg, lam_g0)
    return out['x'], out['status'], out['iters']


def feas_batch(tb, X0, P, lbg=None, ubg=None, max_steps=30):
    """oracle_feas_batch: (X, max |violation|, steps) of the feasibility phase."""
    from omg_tools_b200.solver.b200 import pack_tables
    lib = _load()
    T, keep = pack_tables(tb)
    X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, tb.n)
    B = X0.shape[0]
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(B, tb.n_par)
    lb = np.ascontiguousarray(tb.lbg if lbg is None else lbg, dtype=np.float64)
    ub = np.ascontiguousarray(tb.ubg if ubg is None else ubg, dtype=np.float64)
    X = np.empty((B, tb.n))
    viol = np.empty(B)
    steps = np.empty(B, dtype=np.int32)
    lib.oracle_feas_batch(C.byref(T), B, X0.ctypes.data, P.ctypes.data, lb.ctypes.data,
                          ub.ctypes.data, 1 if lb.ndim == 1 else 0, int(max_steps),
                          X.ctypes.data, viol.ctypes.data, steps.ctypes.data)
    del keep
    return X, viol, steps


def solve_batch_full(tb, X0, P, threads=0, options=None, lbg=None, ubg=None,
                     lam_g0=None, _feas=True):
    # the struct definitions are data-format declarations shared with the product
    from omg_tools_b200.solver.b200 import pack_tables, _Options
    lib = _load()
    T, keep = pack_tables(tb)
    opt = _Options()
    lib.oracle_default_options(C.byref(opt))
    options = dict(options or {})
    # 'dense' (the checker's default): dense L S L^T in the tables' envelope order;
    # 'sparse': up-looking L D L^T on the minimum-degree structure, i.e. the same linear-algebra
    # work as the product's sparse kernel -- what bench.py's CPU arm runs
    linear_solver = options.pop('linear_solver', 'dense')
    lib.oracle_set_linear_solver(1 if linear_solver == 'sparse' else 0)
    retry_mu = float(options.pop('retry_mu', 0.))     # host-level retry, as B200Solver.solve_batch
    feas_steps = int(options.pop('feas_steps', 30))   # host-level feasibility phase, likewise
    if _feas is False:
        feas_steps = 0
    for k, v in options.items():
        setattr(opt, k, v)
    X0 = np.ascontiguousarray(X0, dtype=np.float64).reshape(-1, tb.n)
    B = X0.shape[0]
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(B, tb.n_par)
    lb = np.ascontiguousarray(tb.lbg if lbg is None else lbg, dtype=np.float64)
    ub = np.ascontiguousarray(tb.ubg if ubg is None else ubg, dtype=np.float64)
    shared = 1 if lb.ndim == 1 else 0
    lam0 = None if lam_g0 is None else np.ascontiguousarray(lam_g0, dtype=np.float64)
    X = np.empty((B, tb.n))
    LAM = np.empty((B, tb.m))
    F = np.empty(B)
    st = np.empty(B, dtype=np.int32)
    it = np.empty(B, dtype=np.int32)
    lib.oracle_solve_batch(C.byref(T), C.byref(opt), B, X0.ctypes.data, P.ctypes.data,
                           lb.ctypes.data, ub.ctypes.data, shared,
                           lam0.ctypes.data if lam0 is not None else None,
                           X.ctypes.data, LAM.ctypes.data, F.ctypes.data,
                           st.ctypes.data, it.ctypes.data, int(threads))
    del keep
    res = {'x': X, 'lam_g': LAM, 'f': F, 'status': st, 'iters': it}
    if feas_steps > 0 and (st == 2).any():
        idx = np.nonzero(st == 2)[0]
        lb_i, ub_i = (lb, ub) if shared else (lb[idx], ub[idx])
        x1, _, _ = feas_batch(tb, X[idx], P[idx], lb_i, ub_i, feas_steps)
        opts2 = dict(options)
        opts2['linear_solver'] = linear_solver
        r2 = solve_batch_full(tb, x1, P[idx], threads, opts2, lb_i, ub_i, None, _feas=False)
        ok = r2['status'] == 0              # as B200Solver.solve_batch: keep the first result otherwise
        for key in ('x', 'lam_g', 'f', 'status'):
            res[key][idx[ok]] = r2[key][ok]
        res['iters'][idx] += r2['iters']
        st = res['status']
    if retry_mu > 0. and (st != 0).any():
        idx = np.nonzero(st != 0)[0]
        opts2 = dict(options)
        opts2['mu_init'] = retry_mu
        opts2['linear_solver'] = linear_solver
        r2 = solve_batch_full(tb, X0[idx], P[idx], threads, opts2,
                              lb if shared else lb[idx], ub if shared else ub[idx],
                              None if lam0 is None else lam0[idx])
        ok = r2['status'] == 0
        for key in ('x', 'lam_g', 'f', 'status'):
            res[key][idx[ok]] = r2[key][ok]
        res['iters'][idx] += r2['iters']
    return res
