#!/bin/bash
# round 2, call ab (1 GPU): the staged (cp.async.bulk) combine forward / backward and gate-dispatch backward row gathers, forced onto
# LOCAL rows through the experiment hook (MOE_PEER_BULK=2) — parity of the new kernels on one GPU before the 2-GPU measurement,
# and the isolated layer with and without them
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
export MOE_B200_LIB=tools/variants/libmoe_hooks.so
MOE_PEER_BULK=2 timeout 500 python -m pytest tests -m gpu -x -q -k "layer_forward_backward or deterministic or ragged or masked_layer or full_size or wrapper_golden or autocast" > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ab_pytest.log
tail -3 gpurun_out/r2ab_pytest.log
for shape in "50432 384 16 1" "25216 768 32 2"; do
  for mode in 0 2; do
    echo "== shape $shape MOE_PEER_BULK=$mode"
    MOE_PEER_BULK=$mode timeout 200 python tools/layer_bench.py $shape 2>&1 | grep -E "ms_fwd_bwd|combine|gate_dispatch|dispatch_fwd|sum of" | cut -c1-260
  done
done 2>&1 | tee gpurun_out/r2ab_layer_ab.log
