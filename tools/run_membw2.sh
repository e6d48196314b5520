#!/bin/bash
cd "$(dirname "$0")"
# the probe binary is not kept in git: build it on first use (nvcc cross-compiles on the CPU box; the binary
# travels to the GPU box with the snapshot)
[ -x ./membw2 ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o membw2 membw2.cu -lcuda || exit 1
for cfg in "0 4096 14336 4" "0 4096 14336 8" "0 4096 14336 16" "1 4096 14336 4" "1 4096 14336 8" "1 4096 14336 16" "2 4096 14336 128 64" "3 4096 14336 128 64" "2 4096 14336 256 32" "3 4096 14336 256 32" "2 4096 14336 512 16" "2 4096 14336 1024 8" "3 4096 14336 1024 8" "2 4096 14336 128 32" "0 4096 2048 8" "1 4096 2048 8" "2 4096 2048 128 64" "2 4096 2048 512 16" "2 4096 2048 1024 8" "0 14336 2048 8" "1 14336 2048 8" "2 14336 2048 128 64" "2 14336 2048 1024 8"; do
  ./membw2 $cfg
done
