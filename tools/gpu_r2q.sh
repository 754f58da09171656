#!/bin/bash
# round 2 final evidence, one GPU: suite log, benches of every BASELINE config, reference arm,
# ncu full captures (sparse kernel 592-solve launch; XL kernel config 4 at n=406, 148-solve launch), launch list
mkdir -p gpurun_out/r2q
O=gpurun_out/r2q
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_1gpu.json 2> $O/b1.err; tail -c 200 $O/bench_1gpu.json; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_1gpu.json 2> $O/br.err; tail -c 200 $O/bench_ref_1gpu.json; echo
timeout 300 python bench.py --jitter 0.2 > $O/bench_jitter_1gpu.json 2> $O/bj.err
timeout 300 python bench.py --workload config3 --steps 40 --warmup 5 > $O/bench_config3_n1.json 2> $O/c3.err
timeout 300 python bench.py --workload config3 --formations 9 --steps 40 --warmup 5 > $O/bench_config3_f9_n1.json 2> $O/c39.err
timeout 300 python bench.py --workload config4 --steps 3 --warmup 3 > $O/bench_config4_1gpu.json 2> $O/c4.err
timeout 300 python bench.py --workload config4_5obs --steps 3 --warmup 3 > $O/bench_config4_5obs_1gpu.json 2> $O/c45.err
timeout 300 python bench.py --workload config5 --steps 3 --warmup 3 > $O/bench_config5_1gpu.json 2> $O/c5.err
timeout 300 python bench.py --workload config1 --steps 5 --warmup 3 > $O/bench_config1_1gpu.json 2> $O/c1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 3 -c 1 -f \
    -o $O/prof_sp python bench.py --steps 1 --warmup 3 --batch 592 --cpu-sample 1 > $O/ncu_sp.log 2>&1
tail -2 $O/ncu_sp.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:omg_ipm -s 1 -c 1 -f \
    -o $O/prof_xl python bench.py --workload config4_5obs --steps 1 --warmup 1 --batch 148 --cpu-sample 1 > $O/ncu_xl.log 2>&1
tail -2 $O/ncu_xl.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --cpu-sample 1 > $O/launches.log 2>&1
tail -4 $O/launches.csv
timeout 120 python tools/gpu_debug.py config2 2 2>&1 | tail -22 > $O/phases.txt; cat $O/phases.txt
