#!/bin/bash
# round 2, call f: candidate-subset exact recompute in the tcgen05 gate, identity-MMA gate/dispatch backward, single-pass gate wgrad
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1" "262144 384 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 2>&1 | grep -E "gate|scan|combine|dispatch|total|rc=" | tee gpurun_out/r2f_layer_times_$tag.log
done
