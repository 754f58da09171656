#!/bin/bash
# 2-GPU: formation ADMM through the native NCCL exchange (+ check vs single GPU), config3 bench at N=1,2, batch scaling N=2
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "admm or formation" --tb=short 2>&1 | tail -5 > $O/pytest_admm.log; cat $O/pytest_admm.log
timeout 300 python bench.py --workload config3 --steps 40 --warmup 5 > $O/bench_config3_n1.json 2> $O/c3n1.err; cat $O/bench_config3_n1.json; tail -3 $O/c3n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --workload config3 --steps 40 --warmup 5 > $O/bench_config3_n2.json 2> $O/c3n2.err; cat $O/bench_config3_n2.json; tail -5 $O/c3n2.err
NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 examples/formation_admm_multi_gpu.py --agents 64 --iters 20 --check > $O/admm_check_2gpu.json 2> $O/admm_check.err; cat $O/admm_check_2gpu.json; tail -5 $O/admm_check.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bn2.err; tail -c 700 $O/bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 --scaling strong > $O/bench_n2_strong.json 2> $O/bn2s.err; tail -c 700 $O/bench_n2_strong.json
