#!/bin/bash
# round 2, call e: tcgen05 gate forward (persistent, warp-specialised) + tensor-core gate/dispatch backward: tests, device times,
# ncu --set full of the gate/dispatch backward at the config-2 layer shape
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 2>&1 | grep -E "gate|scan|total|rc=" | tee gpurun_out/r2e_layer_times_$tag.log
done
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gate_dispatch_bwd_mma|gate_fwd_umma' \
    -o gpurun_out/r2e_gdb_c2 -f python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"
