#!/bin/bash
# round 1f, call G: staged gate forward: parity + layer bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- layer staged"; python tools/layer_bench.py 2>&1 | grep "gate_fwd\|ms_fwd_bwd" | cut -c1-200
echo "--- layer unstaged"; MOE_GATE_NO_STAGE=1 python tools/layer_bench.py 2>&1 | grep "gate_fwd\|ms_fwd_bwd" | cut -c1-200
echo "--- sweep point d768 E32 k2"; python tools/layer_bench.py 65536 768 32 2 1.25 2>&1 | tail -20 | cut -c1-200
MOE_GATE_NO_STAGE=1 MOE_GDB_NO_STAGE=1 python tools/layer_bench.py 65536 768 32 2 1.25 2>&1 | grep "gate_fwd\|gate_dispatch\|ms_fwd_bwd" | cut -c1-200
