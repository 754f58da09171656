#!/bin/bash
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
OMG_B200_SNW=1 timeout 300 python bench.py --jitter 0.2 > $O/bench_jitter_snw1.json 2> $O/bj1.err
OMG_B200_KERNEL=envelope timeout 300 python bench.py --jitter 0.2 > $O/bench_jitter_env.json 2> $O/bje.err
python - <<PY
import json
for f in ('bench_jitter_snw1','bench_jitter_env'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['stats']['mean_ip_iterations'], d['cpu_baseline']['max_abs_dx_vs_gpu'], d['cpu_baseline']['iterations_equal_frac'])
PY
