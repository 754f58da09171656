#!/bin/bash
# sparse kernel iteration: bench (identical + jittered) and phase timers
mkdir -p gpurun_out/sp3
O=gpurun_out/sp3
export OMG_B200_VERBOSE=1
one() {
  lbl=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 16 "$@" 2> $O/bench_$lbl.err | tail -1 > $O/bench_$lbl.json
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$lbl.json').read())
    print('$lbl', 'solves/s %.0f'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'ctas', d['roofline']['ctas_per_sm'], 'smem', d['roofline']['smem_bytes'], 'e2e %.0f'%d['e2e']['value'], 'iters', d['stats']['mean_ip_iterations'], 'ok', d['stats']['succeeded_frac'])
except Exception as e:
    print('$lbl failed', e); print(open('$O/bench_$lbl.err').read()[-1500:])
PY
}
one ident
one jit --jitter 0.1
timeout 200 python tools/gpu_debug.py config2 4 2>&1 | tail -18 > $O/phases_alone.txt; cat $O/phases_alone.txt
timeout 200 python tools/gpu_debug.py config2 444 2>&1 | tail -15 > $O/phases_loaded.txt; tail -15 $O/phases_loaded.txt
