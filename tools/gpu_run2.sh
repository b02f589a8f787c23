mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "3x" --timeout=240 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_2_pytest_3x_tmem.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -80 > gpurun_out/r02_2_pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r02_2_breakdown.json > gpurun_out/r02_2_bench_default.json 2> gpurun_out/r02_2_bench_default.err
SMK_ENC_PAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_2_bench_nopair.json 2> gpurun_out/r02_2_bench_nopair.err
SMK_X3_TMEM=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_2_bench_x3_smem.json 2> gpurun_out/r02_2_bench_x3_smem.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --precision tf32 > gpurun_out/r02_2_bench_tf32.json 2> gpurun_out/r02_2_bench_tf32.err
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_2_layers_c3_b256.json > gpurun_out/r02_2_layers_c3_b256.txt 2>&1
timeout 600 python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 5 --out gpurun_out/r02_2_layers_c2_b32.json > gpurun_out/r02_2_layers_c2_b32.txt 2>&1
tail -3 gpurun_out/r02_2_pytest_3x_tmem.log gpurun_out/r02_2_pytest_all.log
python - <<'PY'
import json
for f in ('r02_2_bench_default','r02_2_bench_nopair','r02_2_bench_x3_smem','r02_2_bench_tf32'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d.get('parity') and d['parity'].get('vertices_rel'), d.get('full_cycle') and round(d['full_cycle']['value']), d['config']['execution'][:40])
    except Exception as e: print(f,'ERR',e)
PY
