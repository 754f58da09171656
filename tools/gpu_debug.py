"""Side-by-side iteration trace of the CUDA solver and the numpy twin."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from omg_tools_b200 import scenarios as sc
from omg_tools_b200.solver.b200 import B200Solver
from oracle import ipm_ref

name = sys.argv[1] if len(sys.argv) > 1 else 'config1'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pr = getattr(sc, name)(build_solver=False)
tb = pr.father.tables
print(tb.summary())
X0, P = sc.instance_data(pr, B, jitter=0.2, seed=1)
slv = B200Solver(tb, {'trace': 1})
print(slv.info())
t0 = time.time()
res = slv.solve_batch(X0, P)
print('gpu solve_batch wall %.3fs kernel %.3f ms' % (time.time() - t0, slv.last_timing()[0]))
print('status', res['status'], 'iters', res['iters'], 'f', res['f'])
tr = slv.trace(80)
ref = ipm_ref.solve(tb, X0[0], P[0], trace=True)
print('twin', ref.return_status, ref.iters, ref.f)
for k in range(min(len(ref.log), int(res['iters'][0]) + 2, 80)):
    l = ref.log[k]
    print('%3d | gpu f=%.8f c=%.3e d=%.3e mu=%.2e E=%.3e a=%.3e dw=%.1e | twin f=%.8f c=%.3e d=%.3e mu=%.2e E=%.3e' % (
        k, tr[k,1], tr[k,2], tr[k,3], tr[k,4], tr[k,5], tr[k,6], tr[k,7], l[1], l[2], l[3], l[4], l[5]))
print('max|x_gpu - x_twin| =', np.abs(res['x'][0] - ref.x).max(), ' lam:', np.abs(res['lam_g'][0] - ref.lam_g).max())
for b in range(1, min(B, 4)):
    r = ipm_ref.solve(tb, X0[b], P[b])
    print(b, 'twin', r.return_status, r.iters, 'gpu', res['status'][b], res['iters'][b], 'dx', np.abs(res['x'][b] - r.x).max())
ph = slv.trace(512)[510:512].reshape(-1)
names = ['setup/accept', 'row pass', 'col pass+reduce', 'barrier logic', 'sigma pass', 'zero+H gather',
         'W+border+rhs', 'factor:diag', 'factor:panel', 'factor:trailing', 'back solve', 'step pass',
         'line search', 'tail']
tot = ph[:8].sum() + ph[10:14].sum()
names[7:10] = ['factor (total)', '  of it: gathers', '  of it: root']
names += ['  back: root part', '  factor: panel steps']
print('phase cycles of instance 0 (total %.0f, %d iterations -> %.0f cycles/iter):' % (tot, res['iters'][0], tot / max(1, res['iters'][0])))
for nme, c in zip(names, ph[:16]):
    print('  %-16s %12.0f  %5.1f%%' % (nme, c, 100 * c / tot))
