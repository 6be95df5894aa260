#!/bin/bash
# round 1f, call I: db1 column sums out of the dgelu epilogue: tests + A/B (8 / 12 dgelu epilogue warps, fused / separate db1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L="python tools/layer_bench.py"; F='gemm_dgelu\|colsum_db1\|ms_fwd_bwd'
for rep in 1 2; do
echo "--- 8 warps, fused db1";     $L 2>&1 | grep "$F" | cut -c100-190
echo "--- 8 warps, separate db1";  MOE_DB1_SEPARATE=1 $L 2>&1 | grep "$F" | cut -c100-190
echo "--- 12 warps, fused db1";    MOE_B200_LIB=tools/variants/libmoe_dgelu12.so $L 2>&1 | grep "$F" | cut -c100-190
echo "--- 12 warps, separate db1"; MOE_DB1_SEPARATE=1 MOE_B200_LIB=tools/variants/libmoe_dgelu12.so $L 2>&1 | grep "$F" | cut -c100-190
done
MOE_B200_LIB=tools/variants/libmoe_dgelu12.so timeout 300 python -m pytest tests -m gpu -x -q -k "gemm or dgelu or deterministic" 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -x -q -k "dgelu or deterministic or layer_forward" 2>&1 | tail -2
