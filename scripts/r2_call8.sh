#!/bin/bash
# Round-2 GPU call 8 (one GPU, last of the budget): suite on the final tree, the default bench lines, the tensor-core leaf experiment.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r8.log; "$@" >> gpurun_out/r8.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r8.log; }
run timeout 400 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "bench cfg4 rc=$?" | tee -a gpurun_out/r8.log
run timeout 200 python scripts/leaf_tc_experiment.py 1000000 --exact
timeout 200 ncu --set full --clock-control none -k regex:k_x_leaf_gram_tc -s 2 -c 1 python scripts/leaf_tc_experiment.py 1000000 > gpurun_out/ncu_leaf_tc_r02.txt 2>&1; echo "ncu leaf tc rc=$?" | tee -a gpurun_out/r8.log
timeout 200 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_n1.json 2> gpurun_out/bench_cfg2_n1.err; echo "bench cfg2 rc=$?" | tee -a gpurun_out/r8.log
grep -v "^\s*$" gpurun_out/r8.log | tail -n 40 | cut -c1-1300
