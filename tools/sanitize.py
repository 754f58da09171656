"""Small solves for compute-sanitizer (memcheck / racecheck):
   python tools/sanitize.py [config1|config2|config4|freeT|freeT_warm]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from omg_tools_b200 import scenarios as sc

which = sys.argv[1] if len(sys.argv) > 1 else 'config1'
if which == 'config4':          # XL kernel: intermediates, tape in scratch
    pr = sc.config4()
    X0, P = sc.instance_data(pr, 2, jitter=0.1, seed=3)
    pr.problem.set_options({'max_iter': 3})
elif which.startswith('freeT'):  # soft-restoration path (fires at the first iterations)
    pr = sc.config_freeT()
    X0, P = sc.instance_data(pr, 2, jitter=0.1, seed=2)
    pr.problem.set_options({'max_iter': 12})
elif which == 'config2':        # sparse kernel: supernodes, panel steps, panelised root, early rejection
    pr = sc.config2()
    X0, P = sc.instance_data(pr, 2, jitter=0.2, seed=1)
    pr.problem.set_options({'max_iter': 8})
else:
    pr = sc.config1()
    X0, P = sc.instance_data(pr, 2, jitter=0.2, seed=1)
    pr.problem.set_options({'max_iter': 6})
res = pr.problem.solve_batch(X0, P)
print(which, 'status', res['status'], 'iters', res['iters'])
