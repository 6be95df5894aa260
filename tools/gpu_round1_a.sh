#!/bin/bash
# first GPU measurement pass: bench line, launch list of one step, full ncu capture of the grouped GEMMs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-layer --profile-window"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:grouped_gemm -c 20 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
