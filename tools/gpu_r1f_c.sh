#!/bin/bash
# round 1f, call C: half-slab / 12-warp fp32 epilogue
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; timeout 120 python tools/gemm_probe.py "$@" 2>&1 | tail -3; echo "rc=${PIPESTATUS[0]}"; }
{
run 4 2 128 64 64 0
run 4 2 128 128 128 0
run 4 4 300,5,0,129 192 768 0
run 4 3 640,64,1 1536 384 0
run 4 4 300,5,0,129 768 192 0
run 4 2 1024 384 1536 0
run 5 2 128 64 64 0
run 5 2 1024 1536 384 0
run 5 3 700,0,129 3072 768 0
run 5 16 3152 1536 384 0
MOE_WGRAD_NO_SPLIT=1 run 4 2 1024 1536 384 0
MOE_WGRAD_BN=192 run 5 4 300,5,0,129 768 192 0
} > gpurun_out/probes.log 2>&1
grep -c PASS gpurun_out/probes.log; grep -B4 "FAIL\|rc=[1-9]\|False" gpurun_out/probes.log | tail -30
GB="python tools/gemm_bench.py --E 16 --rows 3152 --iters 30 --no-cublas"
echo "--- d384 new"; $GB --d 384 --ops wgrad1,wgrad2 2>&1 | tail -3 | tee gpurun_out/gemm_new.log
echo "--- d384 wgrad nosplit"; MOE_WGRAD_NO_SPLIT=1 $GB --d 384 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d384 bn192"; MOE_WGRAD_BN=192 $GB --d 384 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d384 bn256"; MOE_WGRAD_BN=256 $GB --d 384 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d768 new"; $GB --d 768 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d768 bn256"; MOE_WGRAD_BN=256 $GB --d 768 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d1024"; $GB --d 1024 --E 8 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
