#!/bin/bash
# single GPU: bench + phase counters of the sparse kernel (config2) + sparse-kernel parity tests
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 300 python bench.py > $O/bench_1gpu.json 2> $O/b1.err; tail -c 300 $O/bench_1gpu.json; echo
timeout 300 python tools/gpu_debug.py config2 2 2>&1 | tail -24 > $O/phases.txt; cat $O/phases.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_parity.log 2>&1; tail -3 $O/pytest_parity.log
