#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --iters 1 --no-cublas --ops fc1,dgelu"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grouped_gemm -c 8 -o gpurun_out/prof_gemm4 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
