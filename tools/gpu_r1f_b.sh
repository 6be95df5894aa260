#!/bin/bash
# round 1f, call B: 384-wide single-accumulator tiles (fc2 / dgrad / wgrad) and split-K weight gradients
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; timeout 120 python tools/gemm_probe.py "$@" 2>&1 | tail -3; echo "rc=${PIPESTATUS[0]}"; }
{
run 1 4 700,300,5,129 0 384 1536
run 1 3 256 0 768 768
run 3 4 300,5,0,129 0 384 1536
run 3 2 1024 0 768 3072
run 4 3 640,64,1 1536 384 0
run 4 2 1024 1536 384 0
run 4 4 300,5,0,129 768 192 0
run 5 2 1024 1536 384 0
run 5 3 700,0,129 3072 768 0
run 5 16 3152 1536 384 0
MOE_WGRAD_NO_SPLIT=1 run 4 2 1024 1536 384 0
} > gpurun_out/probes.log 2>&1
grep -c PASS gpurun_out/probes.log; grep -B4 "FAIL\|rc=[1-9]\|False" gpurun_out/probes.log | tail -30
GB="python tools/gemm_bench.py --E 16 --rows 3152 --iters 30 --no-cublas"
echo "--- d384 new"; $GB --d 384 --ops fc2,dgrad,wgrad1,wgrad2 2>&1 | tail -5 | tee gpurun_out/gemm_new.log
echo "--- d384 wgrad nosplit"; MOE_WGRAD_NO_SPLIT=1 $GB --d 384 --ops wgrad1,wgrad2 2>&1 | tail -3
echo "--- d384 old tiles"; MOE_ROWS_BN=192 MOE_WGRAD_BN=256 $GB --d 384 --ops fc2,dgrad,wgrad1,wgrad2 2>&1 | tail -5
echo "--- d768 new"; $GB --d 768 --ops fc2,dgrad,wgrad1,wgrad2 2>&1 | tail -5
echo "--- d768 old"; MOE_ROWS_BN=256 MOE_WGRAD_BN=256 $GB --d 768 --ops fc2,dgrad,wgrad1,wgrad2 2>&1 | tail -5
echo "--- d192"; $GB --d 192 --E 8 --rows 6304 --ops fc2,dgrad,wgrad1,wgrad2 2>&1 | tail -5
echo "--- layer"; python tools/layer_bench.py 2>&1 | tail -18 | tee gpurun_out/layer_new.log
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
