#!/bin/bash
timeout 100 python bench.py --legs 0 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench $?"; tail -2 gpurun_out/r2g_bench.err
python -c "
import json; b=json.load(open('gpurun_out/r2g_bench.json')); print(b['value'], b['clocks']['samples'], b['config'].get('output_check'))"
