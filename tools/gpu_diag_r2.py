"""Round-2 diagnosis of the five GPU parity failures of round 1 (run on a GPU box):
per-instance status / iteration count / error tables for GPU vs C oracle.
    gpurun -- 'python tools/gpu_diag_r2.py > gpurun_out/diag_r2.log 2>&1'
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omg_tools_b200 import scenarios as sc          # noqa: E402
from oracle import ipm_c                            # noqa: E402

np.set_printoptions(linewidth=200, precision=3)


def cmp(name, pr, X0, P, ncol, threads=8):
    tb = pr.father.tables
    res = pr.problem.solve_batch(X0, P)
    ref = ipm_c.solve_batch_full(tb, X0, P, threads=threads)
    print('==', name, 'n', tb.n, 'm', tb.m)
    print(' status gpu', res['status'], 'ref', ref['status'])
    print(' iters  gpu', res['iters'], 'ref', ref['iters'])
    print(' max|dx| first %d cols' % ncol, np.abs(res['x'] - ref['x'])[:, :ncol].max(axis=1))
    print(' max|dx| all', np.abs(res['x'] - ref['x']).max(axis=1))
    print(' |df|', np.abs(res['f'] - ref['f']))
    sys.stdout.flush()


pr = sc.config_dubins_plain()
X0, P = sc.instance_data(pr, 8, jitter=0.1, seed=1)
cmp('dubins_plain', pr, X0, P, 26)

pr = sc.config_holonomic_orient()
X0, P = sc.instance_data(pr, 4, jitter=0.05, seed=2)
cmp('holonomic_orient', pr, X0, P, 39, threads=4)

pr = sc.config_bicycle()
X0, P = sc.instance_data(pr, 4, jitter=0.02, seed=3)
X0[:, :7] = 0.3
cmp('bicycle', pr, X0, P, 14, threads=4)

# feasibility kernel
for name, seed in (('config5', 5), ('config_dubins_freeT', 3)):
    pr = getattr(sc, name)()
    tb = pr.father.tables
    X0, P = sc.instance_data(pr, 16, jitter=0.2, seed=seed)
    xg, vg, kg = pr.problem.feasibility_batch(X0, P)
    xc, vc, kc = ipm_c.feas_batch(tb, X0, P)
    print('== feas', name)
    print(' steps gpu', kg, 'ref', kc)
    print(' viol gpu', vg, '\n viol ref', vc)
    print(' max|dx|', np.abs(xg - xc).max(axis=1))
    sys.stdout.flush()

# rendezvous, iteration by iteration
from omg_tools_b200.problems.admm_gpu import FormationADMMRunner   # noqa: E402
from oracle.admm_ref import ADMMOracle                              # noqa: E402
run = FormationADMMRunner(sc.config_rendezvous(4))
orc = ADMMOracle(sc.config_rendezvous(4, build_solver=False))
print('== rendezvous')
for it in range(8):
    rg = run.dual_update(0.)
    ro = orc.dual_update(0.)
    st, itg = run.status()
    print(' it', it, 'status gpu', st, 'iters gpu', itg, 'status ref', orc.status,
          'iters ref', getattr(orc, 'iters', None))
    print('   max|dx_i| per agent', np.abs(run.x_i.cpu().numpy() - orc.x_i).max(axis=1))
    print('   max|dz_i|', np.abs(run.z_i.cpu().numpy() - orc.z_i).max(axis=1),
          'max|dl_i|', np.abs(run.l_i.cpu().numpy() - orc.l_i).max(axis=1))
    print('   res gpu', rg, 'ref', ro)
    print('   full x-update max|dX| per agent', np.abs(run.X.cpu().numpy() - orc.X).max(axis=1))
    sys.stdout.flush()
