#!/bin/bash
# First GPU commands of round 2 (run under gpurun from the repo root; everything lands in gpurun_out/).
#   gpurun --timeout 2400 -- 'bash scripts/r2_measure.sh all'          # or one of: tests switches sanitize ab bench cfg3 cfg5 ncu | sharded (2 GPUs)
# Every step runs under its own `timeout` so that a hung experimental kernel cannot hold the box to gpurun's limit.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${1:-all}"
run() { echo "=== $*" | tee -a gpurun_out/r2.log; "$@" >> gpurun_out/r2.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r2.log; }

if [ "$what" = all ] || [ "$what" = tests ]; then
  # the suite that was green at the end of round 1 first, then the files added without a GPU
  run timeout 900 python -m pytest tests -m gpu -x -q -n 4 --deselect tests/test_sharded_gpu.py \
      --ignore tests/test_update.py --ignore tests/test_transformer.py --ignore tests/test_zfull_size_gpu.py
  run timeout 600 python -m pytest tests/test_update.py tests/test_transformer.py tests/test_zfull_size_gpu.py -m gpu -q
  # prepare() with the device-side search-graph bookkeeping (k_search_graph.cu): same structure / query tests, then timing
  NND_PREPARE_GPU=1 run timeout 600 python -m pytest tests/test_gpu_query.py -m gpu -q
  run timeout 600 python scripts/query_bench.py 1000000 128 15 100000
  NND_PREPARE_GPU=1 run timeout 600 python scripts/query_bench.py 1000000 128 15 100000
fi
if [ "$what" = all ] || [ "$what" = switches ]; then
  # the whole parity suite again with every schedule switch ON (they are read by nnd_create, so the env covers every test):
  # the reference goldens, the oracle and the full-size checksum must all still hold
  NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 NND_JOIN_ORDER=leaf NND_CAND_SKIP=1 NND_CAND_KEYS=2 NND_LEAF_SPLIT=1 NND_APPLY_SHORT_SORT=1 \
    NND_APPLY_HALF=1 NND_SMEM_CARVEOUT=1 NND_PREPARE_GPU=1 NND_FOREST_PM=1 run timeout 1500 python -m pytest tests -m gpu -q -n 4 --deselect tests/test_sharded_gpu.py \
    --deselect "tests/test_gpu_parity.py::test_init_random_build_candidates_process"
  # (deselected: the stage test compares old_cand of vertices WITHOUT new candidates, which NND_CAND_SKIP leaves empty by design)
  NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 NND_JOIN_ORDER=leaf NND_CAND_KEYS=1 NND_LEAF_SPLIT=1 NND_APPLY_SHORT_SORT=1 NND_APPLY_HALF=1 NND_FOREST_PM=1 \
    run timeout 600 python -m pytest "tests/test_gpu_parity.py::test_init_random_build_candidates_process" -m gpu -q
fi
if [ "$what" = all ] || [ "$what" = sanitize ]; then
  # compute-sanitizer over one small build with every switch on: out-of-bounds / misaligned accesses (memcheck) and
  # shared-memory hazards (racecheck; mbarrier-ordered accesses may be reported conservatively -- read, do not gate on it)
  ALLSW="NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 NND_JOIN_ORDER=leaf NND_CAND_SKIP=1 NND_CAND_KEYS=2 NND_LEAF_SPLIT=1 NND_APPLY_SHORT_SORT=1 NND_APPLY_HALF=1 NND_PREPARE_GPU=1 NND_FOREST_PM=1"
  run timeout 900 env $ALLSW compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_all.log python scripts/one_build.py 6000 40 10 euclidean 1
  run timeout 900 env $ALLSW compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck_all.log python scripts/one_build.py 3000 40 10 euclidean 1
  run timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_default.log python scripts/one_build.py 6000 40 10 euclidean 1
  tail -n 5 gpurun_out/memcheck_all.log gpurun_out/racecheck_all.log gpurun_out/memcheck_default.log >> gpurun_out/r2.log 2>&1
fi
if [ "$what" = all ] || [ "$what" = ab ]; then
  # schedule experiments: bit-equality + timing, small first (a wrong kernel shows up in seconds), then the bench size
  run timeout 900 python scripts/r2_ab.py 100000 128 15 euclidean --reps=1 --timeout=120
  run timeout 900 python scripts/r2_ab.py 60000 256 15 cosine --reps=1 --timeout=120 --configs=default,rank,leaf,pipe,pipe+rank+leaf,forestpm,jbdk16
  cp gpurun_out/ab_summary.json gpurun_out/ab_summary_small.json 2>/dev/null
  run timeout 1500 python scripts/r2_ab.py 1000000 128 15 euclidean --reps=2 --timeout=200
fi
if [ "$what" = all ] || [ "$what" = bench ]; then
  run timeout 900 python bench.py --steps 5 --warmup 3
  run timeout 600 python bench.py --impl reference --steps 1 --warmup 0
fi
if [ "$what" = all ] || [ "$what" = cfg3 ]; then
  run timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline
  # 1 KB rows: the position-major forest passes gather 370 GB per build here
  NND_FOREST_PM=1 run timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 2 --no-cpu-baseline
fi
if [ "$what" = all ] || [ "$what" = cfg5 ]; then
  # BASELINE.json configs[4]: 1M x 784 correlation k=30 build + query() of 100k held-out points
  run timeout 1500 python scripts/query_bench.py 1000000 784 30 100000 correlation
  # full-size parity of the same build: sha256 of the oracle's 1M x 30 graph (tests/golden/cfg5_full_oracle.json)
  NND_FULL_CFG5=1 run timeout 1500 python -m pytest "tests/test_zfull_size_gpu.py::test_full_size_checksum_of_the_other_configs[cfg5]" -m gpu -q
  run timeout 1200 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline
  # leaf_size 150 / k = 30 does not fit the fused leaf kernel (routed path); the split leaf init does
  NND_LEAF_SPLIT=1 run timeout 1200 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline
  NND_LEAF_SPLIT=1 NND_FOREST_PM=1 run timeout 1200 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline
fi
if [ "$what" = all ] || [ "$what" = ncu ]; then
  # launch list of one warm build (shares), then --set full of the join kernels (default and experimental)
  run timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
      python scripts/one_build.py 1000000 128 15 euclidean 2
  run timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_join_batch -s 17 -c 2 -o gpurun_out/join_batch_r02 -f \
      python scripts/one_build.py 1000000 128 15 euclidean 2
  NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 run timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_join_pipe -s 17 -c 2 \
      -o gpurun_out/join_pipe_r02 -f python scripts/one_build.py 1000000 128 15 euclidean 2
  # the forest's hyperplane pass, position-major (default) and point-major: dram__bytes per level is the number to compare
  run timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forest_side -s 40 -c 6 -o gpurun_out/forest_side_r02 -f \
      python scripts/one_build.py 1000000 128 15 euclidean 2
  NND_FOREST_PM=1 run timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_forest_side -s 40 -c 6 \
      -o gpurun_out/forest_side_pm_r02 -f python scripts/one_build.py 1000000 128 15 euclidean 2
fi
if [ "$what" = sharded ]; then
  # needs `gpurun --gpus 2 -- bash scripts/r2_measure.sh sharded`: the sharded build == the single-GPU build, default schedule and
  # with every switch on (leaf-order join, split leaf init through nnd_leaf_range, candidate skip, half-warp replay ...)
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
  run timeout 900 $TR scripts/sharded_check.py
  NND_JOIN_PIPE=1 NND_RANK_SCATTER=1 NND_JOIN_ORDER=leaf NND_CAND_SKIP=1 NND_CAND_KEYS=2 NND_LEAF_SPLIT=1 NND_APPLY_SHORT_SORT=1 \
    NND_APPLY_HALF=1 NND_SMEM_CARVEOUT=1 NND_FOREST_PM=1 run timeout 900 $TR scripts/sharded_check.py
  run timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline
  NND_LEAF_EXCHANGE=targeted run timeout 900 $TR scripts/sharded_check.py
  NND_LEAF_EXCHANGE=targeted run timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline
fi
tail -n 60 gpurun_out/r2.log
