#!/bin/bash
# After the PDL fix in the grouped GEMV: MoE tests, the bench line (Mixtral leg validates itself against a no-PDL reference)
timeout 200 python -m pytest tests/test_gpu_moe.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench $?"; tail -3 gpurun_out/r2e_bench_n1.err
python -c "
import json; b=json.load(open('gpurun_out/r2e_bench_n1.json')); print(b['value'], b['e2e']['value'], b['roofline']['frac'], b['clocks']); print(json.dumps(b['config']['mixtral'])[-520:])"
