#!/bin/bash
# Round-2 GPU call 10 (one GPU, last minutes): launch list of one 10M build on the FINAL tree + ncu --set full of k_join_batch (raw-page summary).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_10m_final.csv python scripts/one_build.py 10000000 128 15 euclidean 1 > gpurun_out/r10_a.log 2>&1; echo "launch list rc=$?" | tee gpurun_out/r10.log
timeout 240 ncu --set full --clock-control none -k regex:k_join_batch -c 2 -o gpurun_out/join_batch_10m_final -f python scripts/one_build.py 10000000 128 15 euclidean 1 > gpurun_out/r10_b.log 2>&1; echo "ncu full rc=$?" | tee -a gpurun_out/r10.log
python scripts/ncu_summary.py gpurun_out/join_batch_10m_final.ncu-rep > gpurun_out/ncu_join_batch_10m_final.txt 2>&1; echo "summary rc=$?" | tee -a gpurun_out/r10.log
rm -f gpurun_out/join_batch_10m_final.ncu-rep
head -c 1500 gpurun_out/ncu_join_batch_10m_final.txt | tee -a gpurun_out/r10.log
