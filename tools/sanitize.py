"""Small solve for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from omg_tools_b200 import scenarios as sc
pr = sc.config1()
X0, P = sc.instance_data(pr, 2, jitter=0.2, seed=1)
pr.problem.set_options({'max_iter': 6})
res = pr.problem.solve_batch(X0, P)
print('status', res['status'], 'iters', res['iters'])
