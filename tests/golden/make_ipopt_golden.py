"""Record IPOPT's own solutions of the BASELINE configurations -- the fixture that would turn
"parity unpinned" into a pinned parity.

    python tests/golden/make_ipopt_golden.py        ->  tests/golden/ipopt_golden.npz

Needs ``import casadi`` (with its bundled IPOPT), which does NOT work in the authoring image and on
the GPU boxes of this build (no network, no wheel).  Run it wherever CasADi is available; commit
the .npz; tests/test_oracle.py::test_against_ipopt_golden and tests/test_gpu_parity.py pick it up.

What is solved: the lowered tables of this framework (pinned row by row to the reference's own
modelling code by tests/test_model.py) are turned back into CasADi SX expressions -- parameter
tape, polynomial rows, objective -- and handed to ``nlpsol('solver', 'ipopt', ...)`` with exactly
the options the reference sets (omgtools/problems/problem.py:54-62: tol 1e-3,
warm_start_init_point yes, print_level 0, fixed_variable_treatment make_constraint; expand True as
in optilayer.py:55-60), from the same cold starts and parameters the parity tests use
(scenarios.instance_data, seeds below).  Stored per configuration: X0, P, IPOPT's x, lam_g, f,
iteration count and return status.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'ipopt_golden.npz')

CONFIGS = [('config1', 8, 0.1, 1), ('config2', 8, 0.1, 1), ('config5', 8, 0.1, 1),
           ('config4', 4, 0.05, 1), ('config3_agent', 4, 0.1, 1)]


def casadi_nlp(tb, ca):
    """{x, p, f, g} as SX from lowered tables (include/omg_b200.h: omg_tables)."""
    x = ca.SX.sym('x', tb.n)
    p = ca.SX.sym('p', tb.n_par)
    V = [ca.SX(1.)] + [p[k] for k in range(tb.n_par)]
    for e in range(tb.n_tape):                                   # parameter tape, level by level
        acc = ca.SX(0.)
        for t in range(tb.tape_ptr[e], tb.tape_ptr[e + 1]):
            f0, f1, f2, f3 = tb.tape_fac[4 * t:4 * t + 4]
            acc = acc + float(tb.tape_coef[t]) * V[f0] * V[f1] * V[f2] * V[f3]
        fn = int(tb.tape_func[e])
        acc = {0: lambda a: a, 1: lambda a: 1. / a, 2: lambda a: ca.if_else(a >= 0, 1., 0.),
               3: lambda a: ca.if_else(a > 0, 1., 0.), 4: ca.sin, 5: ca.cos, 6: ca.sqrt}[fn](acc)
        V.append(acc)
    xe = [x[j] for j in range(tb.n)] + [ca.SX(1.)]

    def slot(L, s):
        acc = ca.SX(0.)
        for t in range(L.ptr[s], L.ptr[s + 1]):
            term = float(L.coef[t]) * V[int(L.cidx[t])]
            for k in range(L.width):
                term = term * xe[int(L.xi[t * L.width + k])]
            acc = acc + term
        return acc
    for l in range(getattr(tb, 'n_mid', 0)):                     # intermediates: x_ext = [x, 1, mids]
        xe.append(slot(tb.G, tb.m + l))
    g = ca.vertcat(*[slot(tb.G, i) for i in range(tb.m)])
    f = slot(tb.F, 0) if tb.F.n_out == 1 else sum(slot(tb.F, s) for s in range(tb.F.n_out))
    return {'x': x, 'p': p, 'f': f, 'g': g}


def main():
    import casadi as ca
    from omg_tools_b200 import scenarios as sc
    opts = {'expand': True, 'ipopt.tol': 1e-3, 'ipopt.warm_start_init_point': 'yes', 'ipopt.print_level': 0,
            'print_time': 0, 'ipopt.fixed_variable_treatment': 'make_constraint'}
    out = {}
    for name, B, jitter, seed in CONFIGS:
        if name == 'config3_agent':
            pr = sc.config3(4, build_solver=False)
            tb = pr.tb
            X0, P = pr.X[:B].copy(), pr.pack_parameters(0.)[:B]
        else:
            pr = getattr(sc, name)(build_solver=False)
            tb = pr.father.tables
            X0, P = sc.instance_data(pr, B, jitter=jitter, seed=seed)
        solver = ca.nlpsol('solver', 'ipopt', casadi_nlp(tb, ca), opts)
        X, LAM, F, IT, ST = [], [], [], [], []
        for b in range(B):
            r = solver(x0=X0[b], p=P[b], lbg=tb.lbg, ubg=tb.ubg)
            st = solver.stats()
            X.append(np.array(r['x']).reshape(-1)); LAM.append(np.array(r['lam_g']).reshape(-1))
            F.append(float(r['f'])); IT.append(int(st['iter_count'])); ST.append(str(st['return_status']))
        out.update({name + '_X0': X0, name + '_P': P, name + '_x': np.array(X), name + '_lam_g': np.array(LAM),
                    name + '_f': np.array(F), name + '_iters': np.array(IT), name + '_status': np.array(ST)})
        print(name, 'IPOPT:', ST, IT)
    out['casadi_version'] = np.array(ca.__version__)
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    try:
        import casadi  # noqa: F401
    except ImportError:
        sys.exit('casadi is not importable here: nothing recorded (the parity stays unpinned)')
    main()
