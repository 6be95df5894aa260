#!/bin/bash
# round 2, call l: parallel route scan, trimmed gate epilogue, unrolled combine kernels: full GPU suite + device times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -4 gpurun_out/r2l_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 2>&1 | grep -E "gate|scan|combine|dispatch|total|rc=" | tee gpurun_out/r2l_layer_times_$tag.log
done
