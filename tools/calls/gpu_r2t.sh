#!/bin/bash
# round 2, call t (2 GPUs): expert-parallel parity on both transports (with the NaiveGate case on peer memory), then the
# config-2 bench at N = 2 over peer memory (stream-K weight gradients at E_local = 8)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ep.py -m gpu -x -q > gpurun_out/r2t_ep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_ep_pytest.log
tail -4 gpurun_out/r2t_ep_pytest.log; cat gpurun_out/ep_worker_peer_w2.log | grep -E "case|EP_OK"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r2t_bench_c2_n2_peer.json 2> gpurun_out/r2t_bench_c2_n2_peer.err
echo "bench rc=$?"; grep -v -i warn gpurun_out/r2t_bench_c2_n2_peer.err | tail -3
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2t_bench_c2_n2_peer.json").read().strip().splitlines()[-1])
print(j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["per_op_ms"])
print(json.dumps(j.get("expert_parallel")), j.get("parity_check"))
PY
