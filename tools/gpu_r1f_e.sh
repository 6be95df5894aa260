#!/bin/bash
# round 1f, call E: specialised gate/dispatch backward, launch list of one training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "--- layer"; python tools/layer_bench.py 2>&1 | tail -18 | tee gpurun_out/layer_new.log
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer --profile-window"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
