#!/bin/bash
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
for e in 1 0; do
  OMG_B200_EARLY_REJECT=$e timeout 300 python bench.py --jitter 0.2 > $O/bench_jitter_e$e.json 2> $O/bj$e.err
  OMG_B200_EARLY_REJECT=$e timeout 300 python bench.py > $O/bench_e$e.json 2> $O/b$e.err
  python - <<PY
import json
for f in ('bench_jitter_e$e','bench_e$e'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['stats'], d['cpu_baseline']['max_abs_dx_vs_gpu'], d['cpu_baseline']['iterations_equal_frac'])
PY
done
