#!/bin/bash
# sparse kernel: phase timers (one block alone / under full load) + ncu full capture
mkdir -p gpurun_out/sp2
O=gpurun_out/sp2
export OMG_B200_VERBOSE=1
timeout 200 python tools/gpu_debug.py config2 4 2>&1 | tail -22 > $O/phases_alone.txt; cat $O/phases_alone.txt
timeout 200 python tools/gpu_debug.py config2 444 2>&1 | tail -18 > $O/phases_loaded.txt; cat $O/phases_loaded.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o $O/prof_sp python bench.py --steps 1 --warmup 3 --batch 444 --cpu-sample 1 > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
