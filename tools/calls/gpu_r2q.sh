#!/bin/bash
# round 2, call q: stream-K weight gradients — GEMM parity / determinism tests, then A/B against the two-way split-K
# library of the previous commit (tools/variants/libmoe_r2o_splitk.so) inside one call
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped_gemm or bundled or deterministic or ragged or full_size or dense_ffn" > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log; tail -5 gpurun_out/r2q_pytest.log
for shape in "--d 384 --E 16 --rows 3152" "--d 384 --E 2 --rows 25216" "--d 384 --E 4 --rows 12608" "--d 768 --E 8 --rows 6304" "--d 1024 --E 8 --rows 3152"; do
  for lib in new old; do
    echo "== $shape [$lib]"
    if [ $lib = old ]; then export MOE_B200_LIB=tools/variants/libmoe_r2o_splitk.so; else unset MOE_B200_LIB; fi
    timeout 200 python tools/gemm_bench.py $shape --no-cublas --iters 30 --ops wgrad1,wgrad2 2>&1 | grep -E "wgrad|all"
  done
done 2>&1 | tee gpurun_out/r2q_wgrad_ab.log
unset MOE_B200_LIB
