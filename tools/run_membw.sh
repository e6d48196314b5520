#!/bin/bash
# 4096 x 28672 int4 = rowbytes 14336 ; 4096 x 4096 = rowbytes 2048 ; 14336 x 4096 = rowbytes 2048
cd "$(dirname "$0")"
# the probe binary is not kept in git: build it on first use (nvcc cross-compiles on the CPU box; the binary
# travels to the GPU box with the snapshot)
[ -x ./membw ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o membw membw.cu -lcuda || exit 1
for cfg in "4096 14336 128 32 0 16" "4096 14336 128 32 1 16" "4096 14336 256 32 0 16" "4096 14336 256 32 1 16" "4096 14336 512 16 0 16" "4096 14336 512 16 1 16" "4096 14336 1024 8 0 16" "4096 14336 1024 8 1 16" "4096 14336 2048 4 1 16" "4096 14336 7168 2 1 12" "4096 14336 14336 1 1 12" "4096 14336 128 32 0 32" "4096 14336 128 32 1 32" "4096 14336 512 16 1 24" "4096 2048 128 32 0 16" "4096 2048 128 32 1 16" "4096 2048 512 16 1 16" "4096 2048 2048 4 1 16" "14336 2048 128 32 0 16" "14336 2048 2048 4 1 16" "14336 2048 512 16 1 16"; do
  ./membw $cfg 20
done
