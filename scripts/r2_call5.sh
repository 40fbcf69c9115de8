#!/bin/bash
# Round-2 GPU call 5 (one GPU): the suite with the presence / duplicate filters, A/B at 10M, ncu --set full (text kept, reports dropped), bench lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r5.log; "$@" >> gpurun_out/r5.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r5.log; }
ncu_txt() { # name, then the ncu selection, then the command: the details page goes to a text file, no report is kept
  local name=$1; shift
  timeout 600 ncu --set full --clock-control none "$@" > gpurun_out/$name.txt 2>&1
  grep -c "Duration" gpurun_out/$name.txt | sed "s/^/$name kernels profiled: /" | tee -a gpurun_out/r5.log
}
run timeout 900 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py -x
NND_PREPARE_GPU=1 run timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q
run timeout 1500 python scripts/r2_ab.py 10000000 128 15 euclidean --reps=1 --timeout=200 --skip=auto --configs=default,nopresence,nocandskip,jbdk64,jbdk128,forestpm200,forestpm1000
cp gpurun_out/ab_summary.json gpurun_out/ab_summary_10m.json 2>/dev/null
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "bench cfg4 rc=$?" | tee -a gpurun_out/r5.log
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2_n1.json 2> gpurun_out/bench_cfg2_n1.err; echo "bench cfg2 rc=$?" | tee -a gpurun_out/r5.log
ncu_txt ncu_join_batch_10m_r02 -k regex:k_join_batch -c 2 python scripts/one_build.py 10000000 128 15 euclidean 1
ncu_txt ncu_descent_kernels_1m_r02 -k "regex:k_cand_fill|k_cand_select|k_scatter_rank|k_apply|k_cand_count" -c 6 python scripts/one_build.py 1000000 128 15 euclidean 1
ncu_txt ncu_forest_1m_r02 -k "regex:k_forest_side|k_forest_scatter" -s 16 -c 4 python scripts/one_build.py 1000000 128 15 euclidean 1
ncu_txt ncu_leaf_1m_r02 -k "regex:^k_join$|k_leaf_push" -s 4 -c 2 python scripts/one_build.py 1000000 128 15 euclidean 1
rm -f /dev/shm/nnd_ab_*.npy
du -sh gpurun_out | tee -a gpurun_out/r5.log
grep -v "^\s*$" gpurun_out/r5.log | tail -n 45 | cut -c1-1300
