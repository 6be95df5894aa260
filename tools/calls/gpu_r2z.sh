#!/bin/bash
# round 2, call z (2 GPUs): gate/dispatch backward over peer memory — staged gather with cp.async.ca (through L1) against
# the shipped cp.async.cg (L2 only); per-call times of the exchange kernels from bench.py's expert_parallel object
mkdir -p gpurun_out
for v in cg ca cg2 ca2; do
  case $v in ca*) export MOE_B200_LIB=tools/variants/libmoe_r2z_ca.so;; *) unset MOE_B200_LIB;; esac
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29715 bench.py --gpus 2 --steps 10 --warmup 3 \
      --no-layer --no-parity-check > gpurun_out/r2z_bench_$v.json 2> gpurun_out/r2z_bench_$v.err
  python - $v <<'PY'
import json, sys
j=json.loads(open(f"gpurun_out/r2z_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], j["value"], j["ms_per_step"], json.dumps(j["expert_parallel"]["per_call_ms"]))
PY
done 2>&1 | tee gpurun_out/r2z_cpasync_ab.log
