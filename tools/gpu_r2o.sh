#!/bin/bash
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python bench.py --workload config4_5obs --steps 2 --warmup 3 > $O/bench_config4_5obs.json 2> $O/c4.err; tail -c 300 $O/bench_config4_5obs.json; echo
timeout 300 python bench.py --workload config4 --steps 2 --warmup 3 > $O/bench_config4.json 2> $O/c42.err; tail -c 300 $O/bench_config4.json; echo
