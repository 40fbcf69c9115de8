#!/bin/bash
# Round-2 GPU call 2: suite with the promoted defaults, A/B of the remaining switches with in-library stage timers,
# launch list + ncu --set full at 10M / 1M, the new default bench line (cfg4 = 10M x 128 on one GPU), CPU-side evidence in the background.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r2.log; "$@" >> gpurun_out/r2.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r2.log; }
(nice -n 5 python scripts/cpu_baseline_runs.py sweep 200000 > gpurun_out/cpu_sweep.log 2>&1; nice -n 5 python scripts/cpu_baseline_runs.py full cfg4 16 > gpurun_out/cpu_full.log 2>&1) &
CPU_PID=$!
run timeout 900 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py
run timeout 900 python scripts/r2_ab.py 1000000 128 15 euclidean --reps=1 --timeout=120 --skip=auto
cp gpurun_out/ab_summary.json gpurun_out/ab_summary_1m_b.json 2>/dev/null
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_10m.csv python scripts/one_build.py 10000000 128 15 euclidean 1
run timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_join_batch -c 2 -o gpurun_out/join_batch_10m_r02 -f python scripts/one_build.py 10000000 128 15 euclidean 1
run timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_leaf_push|k_cand_fill|k_cand_select|k_scatter_rank|k_apply|k_cand_count" -s 10 -c 16 -o gpurun_out/descent_kernels_1m_r02 -f python scripts/one_build.py 1000000 128 15 euclidean 1
run timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_forest_side|k_forest_scatter" -s 16 -c 6 -o gpurun_out/forest_1m_r02 -f python scripts/one_build.py 1000000 128 15 euclidean 1
run timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^k_join$" -s 2 -c 2 -o gpurun_out/leaf_dist_1m_r02 -f python scripts/one_build.py 1000000 128 15 euclidean 1
wait $CPU_PID
cat gpurun_out/cpu_sweep.log gpurun_out/cpu_full.log >> gpurun_out/r2.log
run timeout 900 python bench.py --steps 3 --warmup 3
run timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3
grep -v "^\s*$" gpurun_out/r2.log | tail -n 70 | cut -c1-1500
