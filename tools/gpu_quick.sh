#!/bin/bash
# quick perf comparison: bench with 1 and 2 blocks per SM
for c in 1 2; do
  OMG_B200_CTAS=$c timeout 200 python bench.py --steps 5 --warmup 3 --cpu-sample 16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('CTAS=$c', 'solves/s %.0f'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'ctas', d['roofline']['ctas_per_sm'], 'smem', d['roofline']['smem_bytes'], 'e2e %.0f'%d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'iters', d['stats']['mean_ip_iterations'], 'ok', d['stats']['succeeded_frac'])"
done
