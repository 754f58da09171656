#!/bin/bash
# round 2: ncu evidence for the sparse kernel (full capture of one launch + launch list) + full suite
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -8 > $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o $O/prof_sp python bench.py --steps 1 --warmup 3 --batch 592 --cpu-sample 1 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 1 > $O/launches.log 2>&1
tail -5 $O/launches.csv
