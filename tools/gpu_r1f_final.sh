#!/bin/bash
# round 1f, final call: GPU tests, smoke, the bench line (db1 fused into the dgelu epilogue) and the same with db1 as a separate pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "--- smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "--- bench (default)"; python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_final.json
echo "--- bench (db1 separate)"; MOE_DB1_SEPARATE=1 python bench.py --no-cpu-baseline > gpurun_out/bench_final_sep.json 2> gpurun_out/bench_final_sep.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_final_sep.json
