#!/bin/bash
# round 2, call ae (1 GPU, last GPU seconds of the round): where the ~0.7 ms per step between `e2e` and `value` goes —
# bench.py --e2e-breakdown re-times the K steps device-only, with the staging copy only and with the overlapped prefetch only
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 140 python bench.py --steps 20 --warmup 5 --no-layer --no-parity-check --no-cpu-baseline --e2e-breakdown > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2ae_bench.json").read().strip().splitlines()[-1])
print(round(j["value"], 1), round(j["ms_per_step"], 3), {k: v for k, v in j["e2e"].items() if not k.endswith("readback")})
PY
