#!/bin/bash
# Round-2 GPU call 7 (gpurun --gpus 8): the row-sharded build at 8 ranks == single GPU, then the metric's configuration (10M x 128) at 8 and 4 GPUs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "=== $*" | tee -a gpurun_out/r7.log; "$@" >> gpurun_out/r7.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r7.log; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29582"
run timeout 600 $TR8 scripts/sharded_check.py
timeout 600 $TR8 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n8.json 2> gpurun_out/bench_cfg4_n8.err; echo "bench cfg4 n8 rc=$?" | tee -a gpurun_out/r7.log
timeout 600 $TR4 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4_n4.json 2> gpurun_out/bench_cfg4_n4.err; echo "bench cfg4 n4 rc=$?" | tee -a gpurun_out/r7.log
tail -c 1200 gpurun_out/bench_cfg4_n8.err gpurun_out/bench_cfg4_n4.err >> gpurun_out/r7.log
grep -v "^\s*$" gpurun_out/r7.log | tail -n 40 | cut -c1-1500
