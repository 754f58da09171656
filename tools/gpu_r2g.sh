#!/bin/bash
# single GPU: full GPU suite (new tests), default bench, config3 with 9 side-by-side formations, XL phase counters
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_1gpu.json 2> $O/b1.err; tail -c 300 $O/bench_1gpu.json; echo
timeout 300 python bench.py --workload config3 --formations 9 --steps 40 --warmup 5 > $O/bench_config3_f9.json 2> $O/c9.err; tail -c 600 $O/bench_config3_f9.json; echo
timeout 300 python tools/gpu_xl_phases.py 2 148 > $O/xl_phases_2obs.txt 2>&1; cat $O/xl_phases_2obs.txt
timeout 300 python tools/gpu_xl_phases.py 5 148 > $O/xl_phases_5obs.txt 2>&1; cat $O/xl_phases_5obs.txt
