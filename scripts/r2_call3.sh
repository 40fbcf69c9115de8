#!/bin/bash
# Round-2 GPU call 3 (one GPU): everything call 2 measured but could not bring home (gpurun_out/ > 64 MiB): bench lines to files,
# launch list at 10M, ncu --set full reduced to text on the box (.ncu-rep deleted), A/B with stage timers; plus cfg3 / cfg5.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r3.log; "$@" >> gpurun_out/r3.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r3.log; }
ncu_txt() { # name, then the ncu selection + command; keeps only the text summary
  local name=$1; shift
  timeout 600 ncu --set full --clock-control none "$@" -o gpurun_out/$name -f > gpurun_out/$name.log 2>&1
  python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep; tail -n 3 gpurun_out/$name.log >> gpurun_out/r3.log; rm -f gpurun_out/$name.log
}
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "bench cfg4 rc=$?" | tee -a gpurun_out/r3.log
run timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_10m.csv python scripts/one_build.py 10000000 128 15 euclidean 1
run timeout 900 python scripts/r2_ab.py 1000000 128 15 euclidean --reps=1 --timeout=120 --skip=auto
run timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k row_sharded_candidate
ncu_txt ncu_join_batch_10m_r02 -k regex:k_join_batch -c 2 python scripts/one_build.py 10000000 128 15 euclidean 1
ncu_txt ncu_descent_kernels_1m_r02 -k "regex:k_leaf_push|k_cand_fill|k_cand_select|k_scatter_rank|k_apply|k_cand_count" -s 10 -c 8 python scripts/one_build.py 1000000 128 15 euclidean 1
ncu_txt ncu_forest_1m_r02 -k "regex:k_forest_side|k_forest_scatter" -s 16 -c 4 python scripts/one_build.py 1000000 128 15 euclidean 1
ncu_txt ncu_leaf_dist_1m_r02 -k "regex:^k_join$" -s 2 -c 1 python scripts/one_build.py 1000000 128 15 euclidean 1
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err; echo "bench cfg3 rc=$?" | tee -a gpurun_out/r3.log
timeout 1500 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg5_n1.json 2> gpurun_out/bench_cfg5_n1.err; echo "bench cfg5 rc=$?" | tee -a gpurun_out/r3.log
NND_FULL_CFG5=1 run timeout 1200 python -m pytest "tests/test_zfull_size_gpu.py::test_full_size_checksum_of_the_other_configs[cfg5]" -m gpu -q
run timeout 1500 python scripts/query_bench.py 1000000 784 30 100000 correlation
run timeout 300 python scripts/cpu_baseline_runs.py sweep 200000
du -sh gpurun_out | tee -a gpurun_out/r3.log
grep -v "^\s*$" gpurun_out/r3.log | tail -n 60 | cut -c1-1200
