#!/bin/bash
# Round-2 GPU call 6 (one GPU): full suite with the partitioned candidate build + register-prefetched presence filter, A/B at 10M, bench lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r6.log; "$@" >> gpurun_out/r6.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r6.log; }
run timeout 1200 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py
run timeout 1500 python scripts/r2_ab.py 10000000 128 15 euclidean --reps=1 --timeout=200 --skip=auto --configs=default,nopresence,nocandpartition,nopm4
cp gpurun_out/ab_summary.json gpurun_out/ab_summary_10m_call6.json 2>/dev/null
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg4_n1.json 2> gpurun_out/bench_cfg4_n1.err; echo "bench cfg4 rc=$?" | tee -a gpurun_out/r6.log
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_n1.json 2> gpurun_out/bench_cfg2_n1.err; echo "bench cfg2 rc=$?" | tee -a gpurun_out/r6.log
rm -f /dev/shm/nnd_ab_*.npy
grep -v "^\s*$" gpurun_out/r6.log | tail -n 45 | cut -c1-1300
