"""Generate tests/golden/p2p_golden.npz: seeded instances of BASELINE configs
1, 2, 5 with the solutions of the CPU oracle (oracle/ipm_ref.py) at the
reference's default tolerance (1e-3) and at 1e-8, plus an independent
scipy SLSQP optimum for config 1.

    python tests/golden/make_p2p_golden.py

"parity unpinned": CasADi/IPOPT cannot run in this image, so these vectors pin
the CUDA path to the oracle, and the oracle's optimum to SLSQP -- not to IPOPT.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omg_tools_b200 import scenarios as sc          # noqa: E402
from oracle import ipm_ref                            # noqa: E402
from oracle.nlp_eval import TableEval                 # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'p2p_golden.npz')
TIGHT = {'tol': 1e-8, 'compl_inf_tol': 1e-8, 'constr_viol_tol': 1e-8}


def main():
    out = {}
    for name, B in (('config1', 4), ('config2', 3), ('config5', 2)):
        pr = getattr(sc, name)(build_solver=False)
        tb = pr.father.tables
        X0, P = sc.instance_data(pr, B, jitter=0.2, seed=1)
        out[name + '_X0'], out[name + '_P'] = X0, P
        out[name + '_dims'] = np.array([tb.n, tb.m, tb.n_par])
        for tag, opt in (('loose', {}), ('tight', TIGHT)):
            xs, lams, its, sts, fs = [], [], [], [], []
            for b in range(B):
                r = ipm_ref.solve(tb, X0[b], P[b], options=opt)
                xs.append(r.x); lams.append(r.lam_g); its.append(r.iters)
                sts.append(r.status); fs.append(r.f)
                print(name, tag, b, r.return_status, r.iters, r.f)
            out['%s_%s_x' % (name, tag)] = np.array(xs)
            out['%s_%s_lam' % (name, tag)] = np.array(lams)
            out['%s_%s_iters' % (name, tag)] = np.array(its)
            out['%s_%s_status' % (name, tag)] = np.array(sts)
            out['%s_%s_f' % (name, tag)] = np.array(fs)
        if name == 'config1':
            from scipy.optimize import minimize
            ev = TableEval(tb)
            V = ev.tape(P[0])
            eq = tb.lbg == tb.ubg
            cons = [{'type': 'eq', 'fun': lambda x: ev.g(x, V)[eq],
                     'jac': lambda x: ev.jac_dense(x, V)[eq]},
                    {'type': 'ineq', 'fun': lambda x: -ev.g(x, V)[~eq],
                     'jac': lambda x: -ev.jac_dense(x, V)[~eq]}]
            res = minimize(lambda x: ev.f(x, V), X0[0], jac=lambda x: ev.gradf(x, V),
                           constraints=cons, method='SLSQP',
                           options={'maxiter': 500, 'ftol': 1e-12})
            print('slsqp', res.status, res.fun)
            out['config1_slsqp_x'], out['config1_slsqp_f'] = res.x, res.fun
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
