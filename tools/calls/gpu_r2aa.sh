#!/bin/bash
# round 2, call aa (1 GPU): staged add+LayerNorm backward (cp.async.bulk row ring per warp) against the register kernel —
# GPU tests of the fused norm, the variant sweep in one process (hooks library), ncu --set full of both kernels, and a short
# bench run with the staged kernel (also the first run of the pipelined e2e loss read-back)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "add_layer_norm or fused_block or weight_copies" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
tail -3 gpurun_out/r2aa_pytest.log
export MOE_B200_LIB=tools/variants/libmoe_lnhooks.so
timeout 300 python tools/ln_sweep.py 2>&1 | tee gpurun_out/r2aa_ln_sweep.log
python tools/ln_prof.py > gpurun_out/plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:addln_bwd_ -o gpurun_out/r2aa_prof_ln -f python tools/ln_prof.py > gpurun_out/ncu_ln.log 2>&1
echo "ncu rc=$?"
for v in bulk; do
  case $v in reg) export MOE_LN_BWD_MODE=0;; *) unset MOE_LN_BWD_MODE;; esac
  timeout 300 python bench.py --steps 20 --warmup 5 --no-layer --no-parity-check --no-cpu-baseline > gpurun_out/r2aa_bench_$v.json 2> gpurun_out/r2aa_bench_$v.err; echo "bench $v rc=$?"
  python - $v <<'PY'
import json, sys
j = json.loads(open(f"gpurun_out/r2aa_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(j["value"], 1), round(j["ms_per_step"], 3), {k: (round(v, 2) if isinstance(v, float) else v) for k, v in j["e2e"].items() if "readback" != k and k != "sync_readback"})
PY
done 2>&1 | tee gpurun_out/r2aa_bench_ab.log
ls -la gpurun_out/*.ncu-rep
