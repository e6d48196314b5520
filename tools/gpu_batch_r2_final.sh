timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_moe.py tests/test_gpu_program.py tests/test_gpu_reference_dropin.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4
timeout 300 python tools/m_sweep.py > gpurun_out/r2_m_sweep3.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2_m_sweep3.json'))
for k,v in d['gemm_layout'].items():
    if any(k.endswith(' M=%d'%m) for m in (1,2,4,8)): print(k, v['us'], v['frac_hbm'])
for k,v in d['other_layouts'].items(): print(k, v['us'])"
timeout 100 python tools/eager_overhead.py 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 --legs 0 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['config']['per_op_path'], d['config']['program_calibration'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"program_kernel|gemv_v3_kernel|rmsnorm_kernel|silu_mul_kernel|stream_pack_kernel" -c 3000 --csv --log-file gpurun_out/r2_ncu_launches_bench_decode.csv python bench.py --steps 2 --warmup 3 --legs 0 > gpurun_out/r2_ncu_launches.log 2>&1
grep -c "program_kernel" gpurun_out/r2_ncu_launches_bench_decode.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^program_kernel -s 3 -c 1 -o gpurun_out/r2_splitk_program_kernel_prof python bench.py --steps 2 --warmup 3 --legs 0 > gpurun_out/r2_ncu_prog.log 2>&1
ls -la gpurun_out/r2_splitk_program_kernel_prof.ncu-rep
