#!/bin/bash
# round 2, call d: GPU tests with the reworked gate forward (registers only, producer folded into warp 0) and the tensor-core
# gate/dispatch backward; ncu device times per kernel at three layer shapes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  bash tools/gpu_layer_times.sh $shape bf16 2>&1 | grep -E "gate|total|rc=" | tee gpurun_out/r2d_layer_times_$tag.log
done
