"""GPU vs C oracle on the golden instances of a config: status, iterations, max |dx|."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
ge.build()
from omg_tools_b200 import scenarios as sc
from oracle import ipm_c
G = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'p2p_golden.npz'))
for name in sys.argv[1:] or ['config2']:
    pr = getattr(sc, name)()
    tb = pr.father.tables
    X0, P = G[name + '_X0'], G[name + '_P']
    Xn, Pn = sc.instance_data(pr, 1)
    X0, P = np.vstack([X0, Xn]), np.vstack([P, Pn])
    for mode in (0, 1):
        pr.problem.set_options({'inertia_mode': mode})
        res = pr.problem.solve_batch(X0, P)
        ref = ipm_c.solve_batch_full(tb, X0, P, threads=4, options={'inertia_mode': mode})
        print(name, 'mode', mode, 'gpu', res['status'], res['iters'], 'oracle', ref['status'], ref['iters'],
              'max|dx|', np.abs(res['x'] - ref['x']).max(axis=1), 'df', np.abs(res['f'] - ref['f']))
