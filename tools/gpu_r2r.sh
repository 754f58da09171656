#!/bin/bash
# final multi-GPU evidence: run with gpurun --gpus N (N = 2 or 8)
N=${1:-2}
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 "${@:3}"; }
run $N 29611 --steps 5 --warmup 3 > $O/bench_n$N.json 2> $O/b$N.err; tail -c 300 $O/bench_n$N.json; echo
run $N 29612 --steps 5 --warmup 3 --scaling strong > $O/bench_n${N}_strong.json 2> $O/bs$N.err; tail -c 300 $O/bench_n${N}_strong.json; echo
run $N 29613 --workload config3 --steps 40 --warmup 5 > $O/bench_config3_n$N.json 2> $O/c$N.err; tail -c 300 $O/bench_config3_n$N.json; echo
run $N 29614 --workload config3 --formations $((9*N)) --steps 40 --warmup 5 > $O/bench_config3_f$((9*N))_n$N.json 2> $O/cf$N.err; tail -c 300 $O/bench_config3_f$((9*N))_n$N.json; echo
if [ $N = 8 ]; then
  run 4 29615 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/b4.err; tail -c 300 $O/bench_n4.json; echo
  run $N 29616 --impl reference --steps 2 --warmup 1 > $O/bench_ref_n8.json 2> $O/br8.err; tail -c 200 $O/bench_ref_n8.json; echo
fi
