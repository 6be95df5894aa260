#!/bin/bash
# Kernel experiments: builds tools/variants/libmoe_<name>.so with extra -D flags — on the GEMM translation unit only, or
# (--all) on every translation unit.
#   tools/build_variant.sh timeline -DMOE_DBG_TIMELINE ; MOE_B200_LIB=tools/variants/libmoe_timeline.so python tools/gemm_bench.py
#   tools/build_variant.sh --all nopdl -DMOE_NO_PDL
set -e
cd "$(dirname "$0")/.."
all=0
if [ "$1" = "--all" ]; then all=1; shift; fi
name=$1; shift
src=slim-switch-moe-vit_b200/csrc; bld=slim-switch-moe-vit_b200/build
make -s -C $src > /dev/null
mkdir -p tools/variants
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
if [ $all = 1 ]; then
  objs=""
  for tu in api routing gate_mma gate_bwd_mma ep_peer gemm_launch block_fusion; do
    $NV "$@" -c $src/$tu.cu -o tools/variants/${tu}_$name.o &
    objs="$objs tools/variants/${tu}_$name.o"
  done
  wait
else
  $NV "$@" -c $src/gemm_launch.cu -o tools/variants/gemm_launch_$name.o
  objs="$bld/api.o $bld/routing.o $bld/gate_mma.o $bld/gate_bwd_mma.o $bld/ep_peer.o $bld/block_fusion.o tools/variants/gemm_launch_$name.o"
fi
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/variants/libmoe_$name.so $objs
rm -f tools/variants/*_$name.o
echo built tools/variants/libmoe_$name.so
