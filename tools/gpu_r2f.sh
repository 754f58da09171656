#!/bin/bash
# 8-GPU box: weak/strong scaling of the batch bench at N=4,8 and config3 at N=4,8 (native NCCL exchange)
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
for n in 8 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 5 --warmup 3 > $O/bench_n$n.json 2> $O/b$n.err; tail -c 400 $O/bench_n$n.json; echo
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --workload config3 --steps 40 --warmup 5 > $O/bench_config3_n$n.json 2> $O/c$n.err; tail -c 500 $O/bench_config3_n$n.json; echo
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 bench.py --gpus 8 --steps 5 --warmup 3 --scaling strong > $O/bench_n8_strong.json 2> $O/b8s.err; tail -c 400 $O/bench_n8_strong.json
