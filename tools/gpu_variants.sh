#!/bin/bash
cd "$(dirname "$0")/.."
bash tools/run_probes.sh > gpurun_out/probes.log 2>&1; grep -c PASS gpurun_out/probes.log; grep -B3 "FAIL\|rc=[1-9]" gpurun_out/probes.log | head -30
python tools/gemm_bench.py --d 384 --E 16 --rows 3152 --no-cublas
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
