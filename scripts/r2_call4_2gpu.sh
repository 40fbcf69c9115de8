#!/bin/bash
# Round-2 GPU call 4 (gpurun --gpus 2): the row-sharded build over NCCL == the single-GPU build, then strong scaling of cfg2 / cfg4 at 2 GPUs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=${1:-2}
run() { echo "=== $*" | tee -a gpurun_out/r4.log; "$@" >> gpurun_out/r4.log 2>&1; echo "    rc=$?" | tee -a gpurun_out/r4.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29581"
run timeout 900 $TR scripts/sharded_check.py
timeout 600 $TR bench.py --gpus $NG --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_n$NG.json 2> gpurun_out/bench_cfg2_n$NG.err; echo "bench cfg2 n$NG rc=$?" | tee -a gpurun_out/r4.log
timeout 900 $TR bench.py --gpus $NG --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_n$NG.json 2> gpurun_out/bench_cfg4_n$NG.err; echo "bench cfg4 n$NG rc=$?" | tee -a gpurun_out/r4.log
tail -c 1500 gpurun_out/bench_cfg2_n$NG.err gpurun_out/bench_cfg4_n$NG.err >> gpurun_out/r4.log
grep -v "^\s*$" gpurun_out/r4.log | tail -n 50 | cut -c1-1500
