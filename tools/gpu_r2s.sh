#!/bin/bash
# XL kernel with the chunked H gather + tape in shared memory; dual decomposition on the GPU
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 200 python tools/gpu_xl_phases.py 2 148 > $O/xl_phases_2obs.txt 2>&1; tail -19 $O/xl_phases_2obs.txt
timeout 200 python tools/gpu_xl_phases.py 5 148 > $O/xl_phases_5obs.txt 2>&1; tail -19 $O/xl_phases_5obs.txt
timeout 200 python bench.py --workload config4 --steps 3 --warmup 3 > $O/bench_config4_1gpu.json 2> $O/c4.err; tail -c 250 $O/bench_config4_1gpu.json; echo
timeout 200 python bench.py --workload config4_5obs --steps 3 --warmup 3 > $O/bench_config4_5obs_1gpu.json 2> $O/c45.err; tail -c 250 $O/bench_config4_5obs_1gpu.json; echo
timeout 400 python -m pytest tests -m gpu -q --tb=short -k "quadrotor or dubins or bicycle or dual_decomposition or holonomic_orient or trailer" > $O/pytest_xl.log 2>&1; tail -4 $O/pytest_xl.log
