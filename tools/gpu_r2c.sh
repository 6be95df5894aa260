#!/bin/bash
# round 2, call c: ncu --set full of the tensor-core gate forward (c2 and c3 layer shapes) + the new reference-loop tests
mkdir -p gpurun_out
python -m pytest tests/test_gpu_reference_loop.py -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2"; do
  tag=$(echo $shape | tr ' ' '_')
  ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gate_fwd_mma' \
      -o gpurun_out/r2c_gate_$tag -f python tools/layer_prof.py $shape bf16 > gpurun_out/r2c_ncu_$tag.log 2>&1
  echo "ncu $tag rc=$?"
done
