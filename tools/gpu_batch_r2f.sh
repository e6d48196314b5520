#!/bin/bash
# clock sampler check: two short decode runs without the secondary legs
for i in 1 2; do
  timeout 100 python bench.py --legs 0 > gpurun_out/r2f_bench_$i.json 2> gpurun_out/r2f_bench_$i.err; echo "bench $?"
  python -c "
import json; b=json.load(open('gpurun_out/r2f_bench_$i.json')); print(b['value'], b['clocks'])"
done
