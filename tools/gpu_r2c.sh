#!/bin/bash
# round 2: suite + bench lines (config 2 default, config 3 ADMM, config 4 at n=406) on one GPU
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -15 > $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 600 $O/bench_ref.json
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; cat $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --jitter 0.1 --steps 5 > $O/bench_jitter.json 2> $O/bench_jitter.err; tail -c 900 $O/bench_jitter.json
timeout 600 python bench.py --workload config3 --steps 40 --warmup 5 > $O/bench_config3_1gpu.json 2> $O/bench_config3.err; cat $O/bench_config3_1gpu.json; tail -3 $O/bench_config3.err
timeout 900 python bench.py --workload config4_5obs --batch 512 --steps 2 --warmup 3 --cpu-sample 16 > $O/bench_config4_5obs.json 2> $O/bench_config4_5obs.err; cat $O/bench_config4_5obs.json; tail -3 $O/bench_config4_5obs.err
