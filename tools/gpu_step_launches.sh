#!/bin/bash
# every launch of one training step with its device time (ncu, cold-cache / serialised: compare shares)
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer --profile-window"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
