#!/bin/bash
# Quadrotor3D (config 4) on the GPU: timing of a batch of 512 (2 and 5 obstacles)
mkdir -p gpurun_out
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/q3d_time.log
import time, numpy as np
import __graft_entry__ as ge
ge.build()
from omg_tools_b200 import scenarios as sc
for nobs in (2, 5):
    try:
        t0 = time.time()
        pr = sc.config4(n_obstacles=nobs)
        tb = pr.father.tables
        print('config4 obstacles', nobs, 'n', tb.n, 'm', tb.m, 'N', tb.kkt_n, 'env', tb.env_size, 'build %.1fs' % (time.time() - t0), pr.problem.info())
        B = 512
        X0, P = sc.instance_data(pr, 1)
        X0 = np.repeat(X0, B, 0); P = np.repeat(P, B, 0)
        for rep in range(2):
            res = pr.problem.solve_batch(X0, P)
            ms, _ = pr.problem.last_timing()
        print('  batch', B, 'kernel ms', ms, 'solves/s', B / ms * 1e3, 'iters', res['iters'][:4], 'status', np.bincount(res['status']))
    except Exception as e:
        import traceback; traceback.print_exc()
PY
