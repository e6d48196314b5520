#!/bin/bash
# Final verification batch of round 2: full GPU suite, smoke(), the driver's bench line, the M sweep.  Run from the repo
# root on the GPU box; everything lands in gpurun_out/.
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c_gpu_tests.txt 2>&1; tail -3 gpurun_out/r2c_gpu_tests.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; echo "bench $?"; tail -c 300 gpurun_out/r2c_bench_n1.json
timeout 300 python tools/m_sweep.py > gpurun_out/r2c_m_sweep.json 2>/dev/null; echo "msweep $?"
TCQ_MS=5,8,16,32,64,128 timeout 200 python tools/tcq_sweep.py > gpurun_out/r2c_tcq_sweep.json 2>/dev/null; echo "tcq $?"
timeout 120 python tools/tcq_pdl_check.py > gpurun_out/tcq_pdl_check.log 2>&1; echo "pdl check $?"
timeout 200 python tools/mixtral_check.py > gpurun_out/mixtral_check.json 2>/dev/null; echo "mixtral order check $?"
