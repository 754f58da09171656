#!/bin/bash
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 200 python tools/gpu_xl_phases.py 2 148 > $O/xl_phases_2obs.txt 2>&1; head -2 $O/xl_phases_2obs.txt | cut -c1-200; grep -E "H gather|row pass|W\+border|total" $O/xl_phases_2obs.txt
timeout 200 python tools/gpu_xl_phases.py 5 148 > $O/xl_phases_5obs.txt 2>&1; head -2 $O/xl_phases_5obs.txt | cut -c1-200; grep -E "H gather|row pass|W\+border|total" $O/xl_phases_5obs.txt
