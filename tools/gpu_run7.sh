mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "win or conv_tc" --timeout=300 --timeout-method=thread -rf 2>&1 | tail -15 > gpurun_out/r02_7_pytest_win.log
tail -n 3 gpurun_out/r02_7_pytest_win.log
timeout 300 python tools/bench_win.py 2>&1 | tail -3
timeout 300 python tools/bench_win.py --cin 64 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -8 > gpurun_out/r02_7_pytest_parity.log
tail -n 3 gpurun_out/r02_7_pytest_parity.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_7_bench_default.json 2> gpurun_out/r02_7_bench_default.err
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_7_layers_c3_b256.json > gpurun_out/r02_7_layers_c3_b256.txt 2>&1
head -12 gpurun_out/r02_7_layers_c3_b256.txt | cut -c1-160
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_7_bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['value']), round(d['e2e']['value']), 'FULL', round(d['full_cycle']['value']), round(d['full_cycle']['e2e']['value']), d['full_cycle']['roofline']['in_situ']['tensor_frac_tf32_sustained'])
PY
