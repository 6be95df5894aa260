#!/bin/bash
# round 2, call y (8 GPUs): expert-parallel parity at W = 8 on both transports (with the NaiveGate case on peer memory), then
# config 2 and config 3 at N = 8 over peer memory with the end-of-round kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ep.py -m gpu -x -q > gpurun_out/r2y_ep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_ep_pytest.log
tail -4 gpurun_out/r2y_ep_pytest.log; grep -E "case|EP_OK" gpurun_out/ep_worker_peer_w8.log gpurun_out/ep_worker_nccl_w8.log
for cfg in c2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 8 --steps 20 --warmup 5 \
      --config $cfg --no-layer > gpurun_out/r2y_bench_${cfg}_n8_peer.json 2> gpurun_out/r2y_bench_${cfg}_n8_peer.err
  echo "bench $cfg rc=$?"; grep -v -i "warn\|run_backward" gpurun_out/r2y_bench_${cfg}_n8_peer.err | tail -3
  python - $cfg <<'PY'
import json, sys
j=json.loads(open(f"gpurun_out/r2y_bench_{sys.argv[1]}_n8_peer.json").read().strip().splitlines()[-1])
print(sys.argv[1], j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["per_op_ms"], j["parity_check"]["ranks_ok"])
print(json.dumps(j.get("expert_parallel", {}).get("per_call_ms")))
PY
done
