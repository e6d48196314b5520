#!/bin/bash
# compute-sanitizer memcheck + synccheck over the small-M tensor-core kernel (two small shapes: G = 64 with two groups of
# constants per stage, an odd number of k-step pairs).  Run on the GPU box; logs -> gpurun_out/.
set -u
cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
T="tests/test_gpu_parity.py::test_small_m_tma_staged_kernel[512-256-64] tests/test_gpu_parity.py::test_small_m_tma_staged_kernel[1152-384-128]"
for tool in ${1:-memcheck synccheck}; do
  timeout 400 $CS --tool $tool --error-exitcode 77 --launch-timeout 120 python -m pytest $T -m gpu -q -x -p no:cacheprovider \
      > gpurun_out/sanitizer_tcq_$tool.log 2>&1
  echo "$tool rc=$? $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_tcq_$tool.log | sort | uniq -c | tr '\n' ';')"
  tail -2 gpurun_out/sanitizer_tcq_$tool.log
done
