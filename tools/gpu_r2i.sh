#!/bin/bash
# ncu full capture of one sparse-kernel launch (batch 592 = one wave)
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o $O/prof_sp python bench.py --steps 1 --warmup 3 --batch 592 --cpu-sample 1 > $O/ncu.log 2>&1
tail -2 $O/ncu.log
