#!/bin/bash
# Kernel experiments: builds tools/variants/libmoe_<name>.so with extra -D flags on the GEMM translation unit.
#   tools/build_variant.sh timeline -DMOE_DBG_TIMELINE ; MOE_B200_LIB=tools/variants/libmoe_timeline.so python tools/gemm_bench.py
set -e
cd "$(dirname "$0")/.."
name=$1; shift
src=slim-switch-moe-vit_b200/csrc; bld=slim-switch-moe-vit_b200/build
make -s -C $src > /dev/null
mkdir -p tools/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
     -c $src/gemm_launch.cu -o tools/variants/gemm_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/variants/libmoe_$name.so $bld/api.o $bld/routing.o $bld/gate_mma.o $bld/gate_bwd_mma.o $bld/ep_peer.o $bld/block_fusion.o tools/variants/gemm_$name.o
rm tools/variants/gemm_$name.o
echo built tools/variants/libmoe_$name.so
