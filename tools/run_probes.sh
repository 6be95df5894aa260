#!/bin/bash
# Runs every GEMM probe case in its own process with a timeout; prints a one-line verdict per case.
cd "$(dirname "$0")/.."
run() { echo "== $*"; timeout 120 python tools/gemm_probe.py "$@" 2>&1 | tail -4; echo "rc=${PIPESTATUS[0]}"; }
#    op E counts M N K
run 1 2 128 0 64 64        # fc2-type, smallest: one k-block, BN=64
run 1 2 128 0 128 128
run 1 4 300,5,0,129 0 192 384
run 1 3 256 0 256 1536
run 1 4 700,300,5,129 0 384 1536   # BN=192, several tiles per expert
run 0 4 300,5,0,129 0 256 192
run 0 2 512 0 1536 384
run 3 2 128 0 64 64        # dgrad
run 3 4 300,5,0,129 0 192 768
run 2 4 300,5,0,129 0 256 192
run 2 2 600 0 1536 384
run 4 2 128 64 64 0        # wgrad smallest (N tail inside a 128-wide tile)
run 4 2 128 128 128 0
run 4 4 300,5,0,129 192 768 0
run 4 4 300,5,0,129 768 192 0
run 4 2 1024 384 1536 0    # M = 384: second pair tile has an empty half
run 4 2 1024 1536 384 0
run 4 4 300,5,0,129 256 192 0      # BN = 192: B through 64-byte-swizzle atoms
run 4 3 640,64,1 1536 384 0
run 5 2 128 64 64 0        # wgrad, transposed store
run 5 4 300,5,0,129 768 192 0
run 5 2 1024 1536 384 0
run 5 3 700,0,129 3072 768 0
