#!/bin/bash
mkdir -p gpurun_out/r2v
O=gpurun_out/r2v
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 200 python bench.py --workload config4 --steps 3 --warmup 3 > $O/bench_config4_1gpu.json 2> $O/c4.err; tail -c 120 $O/bench_config4_1gpu.json; echo
