"""ORACLE runner for bench.py's cpu_baseline / --impl reference legs: solves a
bounded sample of instances with the CPU oracle on all host cores (one process
per core).  Uses the C restatement (oracle/_build/libipm_oracle.so) when built,
else the numpy twin.  Never imported by the product path."""
import multiprocessing as mp
import os

import numpy as np

_TB = None


def _init(tb):
    global _TB
    _TB = tb


def _solve_one(args):
    from oracle import ipm_ref
    x0, p = args
    r = ipm_ref.solve(_TB, x0, p)
    return r.x, r.status, r.iters


def run(tb, X0, P, threads, linear_solver='sparse'):
    """linear_solver='sparse': L D L^T on the minimum-degree structure -- the same
    linear-algebra work as the GPU kernel (the dense factorisation is the checker's default
    and ~2x slower on config 2); built with -O3 -march=native (oracle/Makefile)."""
    try:
        from oracle import ipm_c
        if ipm_c.available():
            r = ipm_c.solve_batch_full(tb, X0, P, threads, options={'linear_solver': linear_solver})
            return {'kind': 'port', 'impl': 'C (%s factorisation)' % linear_solver, 'x': r['x'],
                    'status': r['status'], 'iters': r['iters']}
    except ImportError:
        pass
    ctx = mp.get_context('fork')
    with ctx.Pool(min(threads, len(X0)), initializer=_init, initargs=(tb,)) as pool:
        out = pool.map(_solve_one, list(zip(X0, P)))
    return {'kind': 'port', 'impl': 'numpy', 'x': np.array([o[0] for o in out]),
            'status': np.array([o[1] for o in out]),
            'iters': np.array([o[2] for o in out])}
