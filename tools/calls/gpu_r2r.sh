#!/bin/bash
# round 2, call r: stream-K weight gradients with L2 eviction hints (A evict_first, B evict_last): parity, A/B timing of
# split-K (previous commit) / stream-K without hints / stream-K with hints, DRAM bytes of each (ncu), per-unit timeline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "grouped_gemm or bundled or deterministic" > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log; tail -3 gpurun_out/r2r_pytest.log
libs="hint: nohint:tools/variants/libmoe_r2q_nohint.so splitk:tools/variants/libmoe_r2o_splitk.so"
for shape in "--d 384 --E 16 --rows 3152" "--d 384 --E 2 --rows 25216" "--d 384 --E 4 --rows 12608" "--d 768 --E 4 --rows 12608"; do
  for l in $libs; do
    name=${l%%:*}; path=${l#*:}
    echo "== $shape [$name]"
    if [ -n "$path" ]; then export MOE_B200_LIB=$path; else unset MOE_B200_LIB; fi
    timeout 200 python tools/gemm_bench.py $shape --no-cublas --iters 30 --ops wgrad1,wgrad2 2>&1 | grep -E "wgrad|all"
  done
done 2>&1 | tee gpurun_out/r2r_wgrad_ab.log
for shape in "--d 384 --E 16 --rows 3152" "--d 384 --E 2 --rows 25216"; do
  for l in $libs; do
    name=${l%%:*}; path=${l#*:}
    if [ -n "$path" ]; then export MOE_B200_LIB=$path; else unset MOE_B200_LIB; fi
    echo "== dram $shape [$name]"
    timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:grouped_gemm -s 6 -c 2 \
        python tools/gemm_bench.py $shape --no-cublas --iters 2 --ops wgrad1,wgrad2 2>&1 | grep -E "grouped_gemm|dram__bytes|gpu__time"
  done
done 2>&1 | tee gpurun_out/r2r_wgrad_dram.log
export MOE_B200_LIB=tools/variants/libmoe_r2q_timeline.so
timeout 200 python tools/gemm_timeline.py --op wgrad1 --tiles 4 > gpurun_out/r2r_timeline_c2.log 2>&1
timeout 200 python tools/gemm_timeline.py --op wgrad1 --E 2 --rows 25216 --tiles 4 > gpurun_out/r2r_timeline_e2.log 2>&1
unset MOE_B200_LIB
cat gpurun_out/r2r_timeline_c2.log gpurun_out/r2r_timeline_e2.log
