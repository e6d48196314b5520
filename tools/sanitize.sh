#!/bin/bash
# compute-sanitizer passes over the hand-rolled synchronisation of the library (run on the GPU box; logs -> gpurun_out/)
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards; synccheck: barrier misuse.
set -u
cd "$(dirname "$0")/.."
CS=/usr/local/cuda/bin/compute-sanitizer
export B200AWQ_WATCHDOG_S=120   # tests/conftest.py: program-kernel watchdog (kernels run ~100x slower under the tools)
TOOLS="${1:-memcheck racecheck synccheck}"
T="tests/test_gpu_program.py::test_program_matches_oracle_op_by_op tests/test_gpu_program.py::test_stream_program_general_groups_and_small_shapes tests/test_gpu_moe.py"
for tool in $TOOLS; do
  timeout 900 $CS --tool $tool --error-exitcode 77 --launch-timeout 120 python -m pytest $T -m gpu -q -x -p no:cacheprovider \
      > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log) summaries: $(grep 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log | sort | uniq -c | tr '\n' ';')"
  tail -3 gpurun_out/sanitizer_$tool.log
done
