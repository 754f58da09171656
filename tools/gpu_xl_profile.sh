#!/bin/bash
# XL kernel (config 4): bench line + ncu full capture of one launch
mkdir -p gpurun_out
timeout 600 python bench.py --workload config4 --batch 512 --steps 3 --warmup 3 --cpu-sample 128 > gpurun_out/bench_config4.json 2> gpurun_out/bench_config4.err
cat gpurun_out/bench_config4.json; tail -3 gpurun_out/bench_config4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o gpurun_out/prof_xl python bench.py --workload config4 --steps 1 --warmup 3 --batch 148 --cpu-sample 1 > gpurun_out/ncu_xl.log 2>&1
tail -3 gpurun_out/ncu_xl.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
