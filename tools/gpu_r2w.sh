#!/bin/bash
# compute-sanitizer on the final sparse kernel (config 2: supernodal factorisation, inertia retries in the first
# iterations) + a quick parity check of the same build
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
timeout 100 compute-sanitizer --tool racecheck python tools/sanitize.py config2 > $O/racecheck_config2.txt 2>&1; tail -3 $O/racecheck_config2.txt
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config2 or config1 or golden or formation_admm_64" > $O/pytest_quick.log 2>&1; tail -2 $O/pytest_quick.log
