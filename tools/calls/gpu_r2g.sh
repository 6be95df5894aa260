#!/bin/bash
# round 2, call g (2 GPUs): expert-parallel parity over NVLink peer memory and over NCCL slabs; config-2 bench lines at N = 2 with both transports
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2g_topo.log 2>&1
timeout 900 python -m pytest tests/test_ep.py tests/test_gpu_reference_loop.py -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -25 gpurun_out/r2g_pytest.log
for tr in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 \
      --ep-transport $tr --no-layer > gpurun_out/r2g_bench_c2_n2_$tr.json 2> gpurun_out/r2g_bench_c2_n2_$tr.err
  echo "bench $tr rc=$?"; tail -3 gpurun_out/r2g_bench_c2_n2_$tr.err; head -c 600 gpurun_out/r2g_bench_c2_n2_$tr.json; echo
done
