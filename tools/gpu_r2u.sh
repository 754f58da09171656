#!/bin/bash
# final check of the round: full GPU suite, default bench, config 4 benches with the last XL changes
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 900 python -m pytest tests -m gpu -q --tb=short > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 200 python bench.py > $O/bench_1gpu.json 2> $O/b1.err; tail -c 200 $O/bench_1gpu.json; echo
timeout 200 python bench.py --workload config4 --steps 3 --warmup 3 > $O/bench_config4_1gpu.json 2> $O/c4.err
timeout 200 python bench.py --workload config4_5obs --steps 3 --warmup 3 > $O/bench_config4_5obs_1gpu.json 2> $O/c45.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
