#!/bin/bash
# multi-GPU weak-scaling check: same command line the driver uses
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
cat gpurun_out/scale_$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N', 'solves/s %.0f'%d['value'], 'ms/step %.2f'%d['ms_per_step'], 'frac %.4f'%d['roofline']['frac'], 'e2e %.0f'%d['e2e']['value'])"
tail -3 gpurun_out/scale_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -1
