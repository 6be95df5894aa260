#!/bin/bash
# round 2, call n (2 GPUs): config-2 bench at N = 2 over peer memory with the backlogged per-call profile
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 \
    --no-layer > gpurun_out/r2n_bench_c2_n2_peer.json 2> gpurun_out/r2n_bench_c2_n2_peer.err
echo "rc=$?"; grep -v Warn gpurun_out/r2n_bench_c2_n2_peer.err | tail -3
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2n_bench_c2_n2_peer.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["per_op_ms"])
print(json.dumps(j["expert_parallel"]))
PY
