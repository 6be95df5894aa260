#!/bin/bash
# round 2, call ad (1 GPU, end of the round): full GPU suite, default bench line (c2, N = 1), launch list of one training step —
# the evidence set for the library as shipped (staged add+LayerNorm backward, everything else as in call w)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2ad_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.log
tail -4 gpurun_out/r2ad_pytest.log
timeout 600 python bench.py > gpurun_out/r2ad_bench_c2_n1.json 2> gpurun_out/r2ad_bench_c2_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2ad_bench_c2_n1.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], {k: v for k, v in j["e2e"].items() if not k.endswith("readback")}, j["roofline"]["frac"], j["moe_layer"]["ms_fwd_bwd"], j.get("gpu_launches"))
print(j["roofline"].get("per_op_ms"), j["cpu_baseline"]["value"])
PY
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 bash tools/gpu_step_launches.sh; cp gpurun_out/launches.csv gpurun_out/r2ad_launches.csv
python tools/summarize_profiles.py launches gpurun_out/r2ad_launches.csv gpurun_out/r2ad_step_launches.md \
    "Round 2 (end): every launch of one training step (config 2, N = 1)"
head -14 gpurun_out/r2ad_step_launches.md
rm -f gpurun_out/launches.csv
