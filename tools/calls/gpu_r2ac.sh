#!/bin/bash
# round 2, call ac (2 GPUs): rows pulled from the peer GPU with cp.async.bulk (combine forward / backward, gate-dispatch backward) —
# expert-parallel parity on the peer transport with the shipped library (bulk pulls on), then the N = 2 bench with the hooks
# library, pulls through the LSU (MOE_PEER_BULK=0) against bulk requests (=1): per-call times of the exchange kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ep.py -m gpu -x -q -k "peer" > gpurun_out/r2ac_ep_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_ep_pytest.log
tail -3 gpurun_out/r2ac_ep_pytest.log
export MOE_B200_LIB=tools/variants/libmoe_hooks.so
for v in lsu bulk; do
  case $v in lsu) export MOE_PEER_BULK=0;; *) export MOE_PEER_BULK=1;; esac
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --steps 10 --warmup 3 \
      --no-layer --no-parity-check --no-cpu-baseline > gpurun_out/r2ac_bench_$v.json 2> gpurun_out/r2ac_bench_$v.err; echo "bench $v rc=$?"
  python - $v <<'PY'
import json, sys
j=json.loads(open(f"gpurun_out/r2ac_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(j["value"], 1), round(j["ms_per_step"], 3), json.dumps(j["expert_parallel"]["per_call_ms"]))
PY
done 2>&1 | tee gpurun_out/r2ac_pull_ab.log
