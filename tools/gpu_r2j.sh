#!/bin/bash
# sensitivity: supernode width
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
for w in 4 3 2 1; do
  OMG_B200_SNW=$w timeout 300 python bench.py --cpu-sample 1 > $O/bench_snw$w.json 2> $O/b$w.err
  python - <<PY
import json
d=json.loads(open('$O/bench_snw$w.json').read().strip().splitlines()[-1])
print('snw', $w, d['value'], d['ms_per_step'])
PY
done
