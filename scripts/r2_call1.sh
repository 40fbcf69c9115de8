#!/bin/bash
# Round-2 GPU call 1: every schedule switch under the parity suite, the A/B at 1M, one 10M x 128 build on one GPU.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r2.log; "$@" >> gpurun_out/r2.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r2.log; }
nproc | tee -a gpurun_out/r2.log; free -g | tee -a gpurun_out/r2.log
NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 NND_JOIN_ORDER=leaf NND_CAND_SKIP=1 NND_CAND_KEYS=2 NND_LEAF_SPLIT=1 NND_APPLY_SHORT_SORT=1 \
  NND_APPLY_HALF=1 NND_SMEM_CARVEOUT=1 NND_PREPARE_GPU=1 NND_FOREST_PM=1 run timeout 900 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py \
  --deselect "tests/test_gpu_parity.py::test_init_random_build_candidates_process" -x
run timeout 1500 python scripts/r2_ab.py 1000000 128 15 euclidean --reps=2 --timeout=150 \
  --configs=default,rank,leaf,leaf+rank,applyshort,applyhalf,carveout,leafsplit,leafsplit+carveout,candskip,candkeys,candtile,candskip+tile,jbdk16,forestpm,forestpm96,pipe,pipe+rank,pipe+rank+leaf,pipe-s3,pipe-s2c3,pipe-dk64s2,pipe-dk64s3,all
cp gpurun_out/ab_summary.json gpurun_out/ab_summary_1m.json 2>/dev/null
run timeout 900 python scripts/big_build.py 10000000 128 15 euclidean 2 --json=gpurun_out/big_10m_default.json
tail -n 80 gpurun_out/r2.log
