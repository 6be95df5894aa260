#!/bin/bash
# round 2, call af (1 GPU): BASELINE configs[2] (ViT-B, E = 32, top-2 GShard) on one GPU with the staged LayerNorm backward
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 python bench.py --config c3 --steps 10 --warmup 3 --no-layer --no-cpu-baseline > gpurun_out/r2af_bench_c3_n1.json 2> gpurun_out/r2af_bench_c3_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2af_bench_c3_n1.json").read().strip().splitlines()[-1])
print(round(j["value"], 1), round(j["ms_per_step"], 3), round(j["e2e"]["value"], 1), j["roofline"]["frac"], j["parity_check"]["bit_exact"])
PY
