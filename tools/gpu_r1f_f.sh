#!/bin/bash
# round 1f, call F: DenseFFN + prefetching LayerNorm backward: tests, LN microbench, bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "--- ln"; python tools/ln_bench.py 2>&1 | tail -6
echo "--- bench"; python bench.py > gpurun_out/bench_r1f2.json 2> gpurun_out/bench_r1f2.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_r1f2.err; cut -c1-400 gpurun_out/bench_r1f2.json
