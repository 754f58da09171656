"""GPU vs C oracle on config 4: per-instance difference by variable block."""
import numpy as np
import __graft_entry__ as ge
ge.build()
from omg_tools_b200 import scenarios as sc
from oracle import ipm_c
TIGHT = {'tol': 1e-8, 'compl_inf_tol': 1e-8, 'constr_viol_tol': 1e-8}
pr = sc.config4()
tb = pr.father.tables
ent = pr.father._var_struct.entries
X0, P = sc.instance_data(pr, 8, jitter=0.1, seed=3)
for name, opts in (('default', None), ('tight', TIGHT)):
    if opts:
        pr.problem.set_options(opts)
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=8, options=opts)
    print(name, 'status', res['status'], ref['status'])
    print(name, 'iters', res['iters'], ref['iters'])
    for b in range(8):
        d = np.abs(res['x'][b] - ref['x'][b])
        k = int(np.argmax(d))
        blk = [key[1] for key, (o, sz, sh) in ent.items() if o <= k < o + sz]
        print('  inst', b, 'max', d.max(), blk, 'vehicle part', d[:78].max(), 'df', abs(res['f'][b] - ref['f'][b]))
