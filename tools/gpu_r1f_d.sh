#!/bin/bash
# round 1f, call D: regression probes, GPU tests, bench line, ncu of the HBM-side kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; timeout 120 python tools/gemm_probe.py "$@" 2>&1 | tail -3; echo "rc=${PIPESTATUS[0]}"; }
{
run 4 4 100 256 64 0
run 4 3 640,64,1 1536 384 0
run 4 4 300,5,0,129 192 768 0
run 5 4 100 256 64 0
run 5 16 3152 1536 384 0
} > gpurun_out/probes.log 2>&1
grep -c PASS gpurun_out/probes.log; grep -B4 "FAIL\|rc=[1-9]\|False" gpurun_out/probes.log | tail -30
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- bench"; python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_r1f.err; cut -c1-900 gpurun_out/bench_r1f.json
echo "--- ncu hbm kernels"
python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gate_dispatch_bwd|gate_fwd|combine_fwd|combine_bwd|dispatch_fwd" -o gpurun_out/prof_hbm_r1f -f python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/ncu_hbm.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out | tail -8
