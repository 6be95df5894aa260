#!/bin/bash
# round 2, call v: fc1 / dgelu epilogue experiments — 16 epilogue warps (fc1), 12 / 16 (dgelu), degree-5 GELU polynomial
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in base fc1w16 dg12 dg16 both16 deg5 deg5w16 base; do
  if [ $v = base ]; then unset MOE_B200_LIB; else export MOE_B200_LIB=tools/variants/libmoe_r2v_$v.so; fi
  echo "== $v"
  timeout 200 python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --no-cublas --iters 40 --ops fc1,dgelu 2>&1 | grep -E "fc1|dgelu"
  timeout 200 python tools/gemm_bench.py --d 768 --E 32 --rows 1576 --no-cublas --iters 20 --ops fc1,dgelu 2>&1 | grep -E "fc1|dgelu"
done 2>&1 | tee gpurun_out/r2v_epilogue_variants.log
for v in fc1w16 dg16 deg5; do
  MOE_B200_LIB=tools/variants/libmoe_r2v_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped_gemm or bundled or ffn_intermediates or dgelu_slab or layer_forward_backward" 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r2v_variant_tests.log
