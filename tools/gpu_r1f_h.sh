#!/bin/bash
# round 1f, call H: ncu --set full of the six grouped-GEMM ops at the config-2 layer shape (traffic + tensor-pipe evidence)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --iters 1 --no-cublas --ops fc1,fc2,dgelu,dgrad,wgrad1,wgrad2"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:grouped_gemm --launch-skip 0 -c 24 -o gpurun_out/prof_gemm_r1f -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log; cat gpurun_out/plain.log | tail -8
