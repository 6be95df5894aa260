#!/bin/bash
# round 2, call j: ncu --set full (with source) of the gate forward / gate-dispatch backward at two layer shapes
mkdir -p gpurun_out
for shape in "50432 384 16 1" "262144 384 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gate_dispatch_bwd_mma|gate_fwd_umma' \
      -o gpurun_out/r2j_gate_$tag -f python tools/layer_prof.py $shape bf16 > gpurun_out/r2j_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
done
