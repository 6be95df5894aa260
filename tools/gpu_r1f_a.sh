#!/bin/bash
# round 1f, call A: correctness of the 192-wide / transposed weight-gradient tiles and the staged gate/dispatch
# backward, then A/B timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; timeout 120 python tools/gemm_probe.py "$@" 2>&1 | tail -2; echo "rc=${PIPESTATUS[0]}"; }
{
run 4 4 300,5,0,129 256 192 0
run 4 3 640,64,1 1536 384 0
run 4 2 1024 384 1536 0
run 5 2 128 64 64 0
run 5 4 300,5,0,129 768 192 0
run 5 2 1024 1536 384 0
run 5 3 700,0,129 3072 768 0
} > gpurun_out/probes.log 2>&1
grep -c PASS gpurun_out/probes.log; grep -B3 "FAIL\|rc=[1-9]" gpurun_out/probes.log | tail -20
echo "--- gemm_bench new"; python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --iters 30 --no-cublas --ops wgrad1,wgrad2,wgrad2_mn 2>&1 | tail -5 | tee gpurun_out/gemm_new.log
echo "--- gemm_bench BN=256"; MOE_WGRAD_BN=256 python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --iters 30 --no-cublas --ops wgrad1,wgrad2,wgrad2_mn 2>&1 | tail -5 | tee gpurun_out/gemm_bn256.log
echo "--- d192"; python tools/gemm_bench.py --d 192 --E 8 --rows 6304 --iters 30 --no-cublas --ops wgrad1,wgrad2,wgrad2_mn 2>&1 | tail -4
echo "--- layer staged"; python tools/layer_bench.py 2>&1 | tail -18 | tee gpurun_out/layer_new.log
echo "--- layer unstaged"; MOE_GDB_NO_STAGE=1 python tools/layer_bench.py 2>&1 | grep "gate_dispatch_bwd\|ms_fwd_bwd" | cut -c1-300
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
