#!/bin/bash
cd "$(dirname "$0")/.."
python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:gate_fwd' \
    -o gpurun_out/prof_gate -f python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/ncu_layer.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_layer.log
