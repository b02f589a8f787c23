mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_masking.py -q --timeout=300 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_4_pytest.log
SMK_X3_TRUNC=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "3x" --timeout=300 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_4_pytest_trunc.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_4_bench_default.json 2> gpurun_out/r02_4_bench_default.err
SMK_X3_TRUNC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline > gpurun_out/r02_4_bench_trunc.json 2> gpurun_out/r02_4_bench_trunc.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity --slots 6 > gpurun_out/r02_4_bench_slots6.json 2> gpurun_out/r02_4_bench_slots6.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity --slots 2 > gpurun_out/r02_4_bench_slots2.json 2> gpurun_out/r02_4_bench_slots2.err
SMK_CONV3_WIN=0 timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 > gpurun_out/r02_4_layers_c3_b256_nowin.txt 2>&1
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_4_layers_c3_b256.json > gpurun_out/r02_4_layers_c3_b256.txt 2>&1
tail -n 4 gpurun_out/r02_4_pytest.log gpurun_out/r02_4_pytest_trunc.log
python - <<'PY'
import json
for f in ('r02_4_bench_default','r02_4_bench_trunc','r02_4_bench_slots6','r02_4_bench_slots2'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d.get('parity') and d['parity'].get('params_rel'), d.get('full_cycle') and round(d['full_cycle']['value']))
    except Exception as e: print(f,'ERR',e)
PY
grep -E "win|head|K288_N32|K576_N32" gpurun_out/r02_4_layers_c3_b256.txt gpurun_out/r02_4_layers_c3_b256_nowin.txt | cut -c1-200
