#!/bin/bash
# XL kernel after the border ordering: phases at n=238 / n=406, config4_5obs bench, full GPU suite
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 300 python tools/gpu_xl_phases.py 2 148 > $O/xl_phases_2obs.txt 2>&1; cat $O/xl_phases_2obs.txt
timeout 300 python tools/gpu_xl_phases.py 5 148 > $O/xl_phases_5obs.txt 2>&1; cat $O/xl_phases_5obs.txt
timeout 300 python bench.py --workload config4_5obs --steps 2 --warmup 3 > $O/bench_config4_5obs.json 2> $O/c4.err; tail -c 300 $O/bench_config4_5obs.json; echo
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
