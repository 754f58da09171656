#!/bin/bash
# First GPU run of the sparse kernel: smoke, parity subset, bench variants.
mkdir -p gpurun_out/sp1
O=gpurun_out/sp1
export OMG_B200_VERBOSE=1
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -3 $O/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --tb=short > $O/pytest_parity.log 2>&1; tail -15 $O/pytest_parity.log
one() {  # label, env...
  lbl=$1; shift
  extra=""; if [ "$lbl" = "sp128_j" ]; then extra="--jitter 0.1"; fi
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --cpu-sample 16 $extra 2> $O/bench_$lbl.err | tail -1 > $O/bench_$lbl.json
  python - <<PY
import json
try:
    d=json.loads(open('$O/bench_$lbl.json').read())
    print('$lbl', 'solves/s %.0f'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'ctas', d['roofline']['ctas_per_sm'], 'smem', d['roofline']['smem_bytes'], 'e2e %.0f'%d['e2e']['value'], 'iters', d['stats']['mean_ip_iterations'], 'ok', d['stats']['succeeded_frac'])
except Exception as e:
    print('$lbl failed', e); print(open('$O/bench_$lbl.err').read()[-1500:])
PY
}
one sp128 OMG_B200_SP_NT=128
one sp256 OMG_B200_SP_NT=256
one env OMG_B200_KERNEL=envelope
one sp128_j OMG_B200_SP_NT=128
