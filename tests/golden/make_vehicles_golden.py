"""Generate tests/golden/vehicles_golden.npz: seeded instances of BASELINE
config 4 (Quadrotor3D, examples/p2p_3dquadrotor.py) and of the Holonomic3D
example (examples/p2p_holonomic_3d.py, interior start/goal) with the solutions
of the numpy oracle (oracle/ipm_ref.py) at the reference's tolerance (1e-3)
and at 1e-8.

    python tests/golden/make_vehicles_golden.py

"parity unpinned": CasADi/IPOPT cannot run in this image (see DESIGN.md
section 5); these vectors pin the C oracle and the CUDA path to the numpy
oracle.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from omg_tools_b200 import scenarios as sc          # noqa: E402
from oracle import ipm_ref                            # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'vehicles_golden.npz')
TIGHT = {'tol': 1e-8, 'compl_inf_tol': 1e-8, 'constr_viol_tol': 1e-8}


def builders():
    yield 'config4', sc.config4(build_solver=False), 0.1, 3
    yield 'holonomic3d', sc.config_holonomic3d(
        build_solver=False, start=(-1.7, -1.7, -1.7), goal=(1.7, 1.7, -1.7)), 0.1, 1


def main():
    out = {}
    for name, pr, jitter, seed in builders():
        tb = pr.father.tables
        X0, P = sc.instance_data(pr, 2, jitter=jitter, seed=seed)
        out[name + '_X0'], out[name + '_P'] = X0, P
        out[name + '_dims'] = np.array([tb.n, tb.m, tb.n_par])
        for tag, opt in (('loose', {}), ('tight', TIGHT)):
            res = [ipm_ref.solve(tb, X0[b], P[b], options=opt) for b in range(2)]
            for b, r in enumerate(res):
                print(name, tag, b, r.return_status, r.iters, r.f)
            out['%s_%s_x' % (name, tag)] = np.array([r.x for r in res])
            out['%s_%s_lam' % (name, tag)] = np.array([r.lam_g for r in res])
            out['%s_%s_iters' % (name, tag)] = np.array([r.iters for r in res])
            out['%s_%s_status' % (name, tag)] = np.array([r.status for r in res])
            out['%s_%s_f' % (name, tag)] = np.array([r.f for r in res])
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
