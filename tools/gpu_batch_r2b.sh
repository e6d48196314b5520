#!/bin/bash
# Round-2 (second session) verification batch: full GPU suite, M sweeps, the driver's bench line, one ncu capture of the
# small-M kernel.  Run from the repo root on the GPU box; everything lands in gpurun_out/.
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_gpu_tests.txt 2>&1; tail -4 gpurun_out/r2b_gpu_tests.txt
timeout 300 python tools/m_sweep.py > gpurun_out/r2b_m_sweep.json 2>/dev/null; echo "msweep $?"
TCQ_MS=5,8,16,32,64,128 timeout 200 python tools/tcq_sweep.py > gpurun_out/r2b_tcq_sweep.json 2>/dev/null; echo "tcq $?"
timeout 600 python bench.py > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench $?"; tail -c 600 gpurun_out/r2b_bench_n1.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tcq -s 2 -c 1 -f -o gpurun_out/r2b_tcq16_big python tools/ncu_target.py gemm16_big > gpurun_out/ncu_tcq.log 2>&1; echo "ncu $?"
