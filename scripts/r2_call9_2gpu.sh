#!/bin/bash
# Round-2 GPU call 9 (gpurun --gpus 2, the last minutes): the sharded upload of X in the end-to-end path (1 / world over PCIe + NVLink all-gather).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581"
timeout 150 $TR scripts/sharded_check.py 200000 > gpurun_out/r9_check.log 2>&1; echo "sharded_check rc=$?" | tee gpurun_out/r9.log
timeout 200 $TR bench.py --gpus 2 --workload cfg2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; echo "bench cfg2 n2 rc=$?" | tee -a gpurun_out/r9.log
grep -h "sharded==single\|SHARDED_CHECK\|Error\|error" gpurun_out/r9_check.log | cut -c1-400 | tail -5 | tee -a gpurun_out/r9.log
tail -c 600 gpurun_out/bench_cfg2_n2.err | tee -a gpurun_out/r9.log
