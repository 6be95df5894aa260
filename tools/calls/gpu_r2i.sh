#!/bin/bash
# round 2, call i (8 GPUs): expert-parallel parity at W = 8 (peer memory + NCCL), and the expert-parallel configs of
# BASELINE.json: c2 / c3 / c4 at N = 8, c3 at N = 4; NCCL transport at N = 8 (c2) for comparison
mkdir -p gpurun_out
run() {   # run <tag> <nproc> <bench args...>
  tag=$1; n=$2; shift 2
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $n --steps 10 --warmup 3 \
      --no-layer "$@" > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err
  echo "$tag rc=$?"; grep -v Warning gpurun_out/r2i_$tag.err | tail -2; head -c 300 gpurun_out/r2i_$tag.json; echo
}
timeout 600 python -m pytest tests/test_ep.py -m gpu -x -q > gpurun_out/r2i_pytest_ep_w8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_ep_w8.log
tail -4 gpurun_out/r2i_pytest_ep_w8.log
run c2_n8_peer 8 --config c2
run c3_n8_peer 8 --config c3
run c4_n8_peer 8 --config c4
run c2_n8_nccl 8 --config c2 --ep-transport nccl
run c3_n4_peer 4 --config c3
