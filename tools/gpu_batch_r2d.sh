timeout 300 python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench $?"
python -c "
import json; b=json.load(open('gpurun_out/r2d_bench_n1.json')); print(b['value'], b['clocks']); print(json.dumps(b['config']['mixtral'])[-400:])"
timeout 200 python tools/mixtral_check.py > gpurun_out/mixtral_check.json 2> gpurun_out/mixtral_check.err; echo "check $?"; tail -2 gpurun_out/mixtral_check.err
python -c "
import json; d=json.load(open('gpurun_out/mixtral_check.json')); print(d['alone'], d['after_other_legs'])"
timeout 120 python tools/tcq_pdl_check.py > gpurun_out/tcq_pdl_check.log 2>&1; echo "pdl check $?"; cat gpurun_out/tcq_pdl_check.log | tail -16
