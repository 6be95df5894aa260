#!/bin/bash
# round 2, call p: the round's evidence set in one call — full GPU suite, default bench line (c2, N = 1) and the CPU
# reference arm, launch list of one training step, per-kernel device times of four layer shapes, ncu --set full of the
# gate kernels and of the six grouped-GEMM launches at the config-2 layer shape.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/r2p_gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -4 gpurun_out/r2p_pytest.log
timeout 600 python bench.py > gpurun_out/r2p_bench_c2_n1.json 2> gpurun_out/r2p_bench_c2_n1.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_reference_arm.json 2> gpurun_out/r2p_bench_reference_arm.err; echo "ref rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2p_bench_c2_n1.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["e2e"], j["roofline"]["frac"], j["moe_layer"]["ms_fwd_bwd"], j.get("gpu_launches"))
print(j["roofline"].get("per_op_ms"))
PY
# launch list of one training step (same bench command, cold-cache / serialised: shares)
timeout 600 bash tools/gpu_step_launches.sh; cp gpurun_out/launches.csv gpurun_out/r2p_launches.csv
python tools/summarize_profiles.py launches gpurun_out/r2p_launches.csv gpurun_out/r2p_step_launches.md \
    "Round 2: every launch of one training step (config 2, N = 1)"
# per-kernel device times + DRAM bytes of one layer fwd+bwd
for shape in "50432 384 16 1" "25216 768 32 2" "65536 1024 64 1" "262144 384 16 1" "262144 384 64 1"; do
  tag=$(echo $shape | tr ' ' '_')
  timeout 300 bash tools/gpu_layer_times.sh $shape bf16 > gpurun_out/r2p_layer_times_$tag.log 2>&1
  tail -1 gpurun_out/r2p_layer_times_$tag.log
done
# ncu --set full: gate / routing / combine kernels, then the six GEMM launches (config-2 layer shape)
python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:gate_|route_scan|dispatch_fwd|combine_' \
    -o gpurun_out/r2p_prof_hbm -f python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k 'regex:grouped_gemm' \
    -o gpurun_out/r2p_prof_gemm -f python tools/layer_prof.py 50432 384 16 1 bf16 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
python tools/summarize_profiles.py full gpurun_out/r2p_prof_hbm.ncu-rep gpurun_out/r2p_hbm_kernels_ncu.md \
    "Round 2: ncu --set full of the gate / scan / dispatch / combine kernels, config-2 layer shape (T = 50 432, d = 384, E = 16, top-1, bf16)"
python tools/summarize_profiles.py full gpurun_out/r2p_prof_gemm.ncu-rep gpurun_out/r2p_gemm_ncu_full.md \
    "Round 2: ncu --set full of the six grouped-GEMM launches, config-2 layer shape"
ls -la gpurun_out | head -40
