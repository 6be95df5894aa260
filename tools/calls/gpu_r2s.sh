#!/bin/bash
# round 2, call s: programmatic dependent launch on every layer kernel (A/B against the same tree built with -DMOE_NO_PDL),
# weight-gradient schedule policy (stream-K / equal split-K / whole tiles) against the split-K library of the round's start
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -4 gpurun_out/r2s_pytest.log
for l in pdl: nopdl:tools/variants/libmoe_r2s_nopdl.so pdl2: nopdl2:tools/variants/libmoe_r2s_nopdl.so; do
  name=${l%%:*}; path=${l#*:}
  if [ -n "$path" ]; then export MOE_B200_LIB=$path; else unset MOE_B200_LIB; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_$name.json 2> gpurun_out/r2s_bench_$name.err; echo "bench $name rc=$?"
  python - "$name" <<'PY'
import json, sys
j = json.loads(open(f"gpurun_out/r2s_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], "layer", j["moe_layer"]["ms_fwd_bwd"], j["roofline"].get("per_op_ms"))
PY
done 2>&1 | tee gpurun_out/r2s_pdl_ab.log
unset MOE_B200_LIB
for shape in "--d 384 --E 16 --rows 3152" "--d 384 --E 8 --rows 6304" "--d 384 --E 4 --rows 12608" "--d 384 --E 2 --rows 25216" "--d 768 --E 4 --rows 12608" "--d 768 --E 32 --rows 1576"; do
  for l in new: splitk:tools/variants/libmoe_r2o_splitk.so; do
    name=${l%%:*}; path=${l#*:}
    echo "== $shape [$name]"
    if [ -n "$path" ]; then export MOE_B200_LIB=$path; else unset MOE_B200_LIB; fi
    timeout 200 python tools/gemm_bench.py $shape --no-cublas --iters 30 --ops wgrad1,wgrad2 2>&1 | grep -E "wgrad|all"
  done
done 2>&1 | tee gpurun_out/r2s_wgrad_ab.log
unset MOE_B200_LIB
