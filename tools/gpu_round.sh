#!/bin/bash
# One GPU round-trip: parity tests, bench line, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --cpu-sample 1 > gpurun_out/ncu_b.log 2>&1
tail -15 gpurun_out/launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --batch 148 --cpu-sample 1 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
