#!/bin/bash
# single GPU: full GPU suite + benches (default, jitter, config3 x1 / x9, config4_5obs) with the supernodal kernel
mkdir -p gpurun_out/r2k
O=gpurun_out/r2k
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_1gpu.json 2> $O/b1.err; tail -c 200 $O/bench_1gpu.json; echo
timeout 300 python bench.py --jitter 0.2 > $O/bench_jitter.json 2> $O/bj.err; tail -c 200 $O/bench_jitter.json; echo
timeout 300 python bench.py --workload config3 --steps 40 --warmup 5 > $O/bench_config3.json 2> $O/c1.err; tail -c 300 $O/bench_config3.json; echo
timeout 300 python bench.py --workload config3 --formations 9 --steps 40 --warmup 5 > $O/bench_config3_f9.json 2> $O/c9.err; tail -c 300 $O/bench_config3_f9.json; echo
timeout 300 python bench.py --workload config4_5obs --steps 2 --warmup 3 > $O/bench_config4_5obs.json 2> $O/c4.err; tail -c 300 $O/bench_config4_5obs.json; echo
