#!/bin/bash
# round 2, call a: sanity of the round-2 host changes on a B200, legacy mma.sync rate, c2 + c3 (all experts local) bench lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpu.log 2>&1
tools/ubench/mma_sync_rate > gpurun_out/r2a_mma_sync_rate.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "c2 rc=$?" >> gpurun_out/r2a_bench_c2.err
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_c3_n1.json 2> gpurun_out/r2a_bench_c3_n1.err; echo "c3 rc=$?" >> gpurun_out/r2a_bench_c3_n1.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_mma_sync_rate.log; tail -2 gpurun_out/r2a_bench_c2.err gpurun_out/r2a_bench_c3_n1.err
