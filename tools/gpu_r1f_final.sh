#!/bin/bash
# round 1f, final call: GPU tests, smoke and the bench line of the committed state
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "--- bench (default)"; python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_final.json
