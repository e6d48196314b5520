#!/bin/bash
# mma.sync issue-rate probe (profiles/r02_hmma_rate.txt); the binary is not kept in git: built on first use
cd "$(dirname "$0")"
[ -x ./hmma_rate ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o hmma_rate hmma_rate.cu || exit 1
./hmma_rate
