#!/bin/bash
# round 2, call k: two-group gate epilogue, lighter gate/dispatch backward; full GPU suite + device times + bench c2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -4 gpurun_out/r2k_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1" "262144 384 64 1" "262144 384 16 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 2>&1 | grep -E "gate|scan|combine|dispatch|cast|colsum|total|rc=" | tee gpurun_out/r2k_layer_times_$tag.log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "c2 rc=$?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2k_bench_c2.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["moe_layer"]["ms_fwd_bwd"])
print({k:v["ms"] for k,v in j["moe_layer"]["kernels"].items()})
PY
