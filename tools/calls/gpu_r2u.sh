#!/bin/bash
# round 2, call u: dgelu / dgrad read the forward bf16 weights MN-major (no transposed copies, plain weight cast):
# full GPU suite, A/B of the row-mode GEMMs against the K-major library of the previous commit, bench c2, layer times c2 / c3
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
for shape in "--d 384 --E 16 --rows 3152" "--d 768 --E 32 --rows 1576" "--d 1024 --E 8 --rows 8192"; do
  echo "== $shape [mn-major B]"
  timeout 200 python tools/gemm_bench.py $shape --no-cublas --iters 30 --ops fc1,fc2,dgelu,dgrad 2>&1 | grep -E "fc|dg|all"
  echo "== $shape [k-major B, transposed copies]"
  MOE_B200_LIB=tools/variants/libmoe_r2u_kmajor.so timeout 200 python tools/variants/gemm_bench_kmajor.py $shape --no-cublas --iters 30 --ops fc1,fc2,dgelu,dgrad 2>&1 | grep -E "fc|dg|all"
done 2>&1 | tee gpurun_out/r2u_rows_ab.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2u_bench_c2.json 2> gpurun_out/r2u_bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2u_bench_c2.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["e2e"]["value"], j["roofline"]["frac"], "layer", j["moe_layer"]["ms_fwd_bwd"], j["roofline"].get("per_op_ms"))
print({k: v["ms"] for k, v in j["moe_layer"]["kernels"].items()})
PY
for shape in "50432 384 16 1" "25216 768 32 2"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 > gpurun_out/r2u_layer_times_$tag.log 2>&1
  grep -E "cast|total" gpurun_out/r2u_layer_times_$tag.log
done
