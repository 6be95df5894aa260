#!/bin/bash
# round 2, call b: GPU tests after the gate test fix; ncu device times per kernel of one layer fwd+bwd at c2 / c4-like / E=64 shapes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1" "262144 384 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  bash tools/gpu_layer_times.sh $shape bf16 > gpurun_out/r2b_layer_times_$tag.log 2>&1
  cat gpurun_out/r2b_layer_times_$tag.log
done
