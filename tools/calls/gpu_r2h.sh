#!/bin/bash
# round 2, call h (2 GPUs): EP parity (peer + nccl) after the kernel changes, GEMM tests with the S-way split, config-3 bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ep.py tests/test_gpu_parity.py -m gpu -x -q -k "ep_parity or grouped_gemm or deterministic or bundled or layer_forward_backward" > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 \
    --config c3 --no-layer > gpurun_out/r2h_bench_c3_n2_peer.json 2> gpurun_out/r2h_bench_c3_n2_peer.err
echo "bench c3 rc=$?"; tail -3 gpurun_out/r2h_bench_c3_n2_peer.err; head -c 400 gpurun_out/r2h_bench_c3_n2_peer.json; echo
