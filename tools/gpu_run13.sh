mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "xdw" --timeout=400 --timeout-method=thread -rf 2>&1 | tail -6
timeout 300 python tools/bench_xdw.py --x3 --batch 256 --reps 5 > gpurun_out/r02_13_xdw_b256_x3.txt 2>&1; head -12 gpurun_out/r02_13_xdw_b256_x3.txt; tail -1 gpurun_out/r02_13_xdw_b256_x3.txt
timeout 300 python tools/bench_xdw.py --x3 --batch 32 --reps 10 > gpurun_out/r02_13_xdw_b32_x3.txt 2>&1; tail -1 gpurun_out/r02_13_xdw_b32_x3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_13_bench.json 2> gpurun_out/r02_13_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_13_bench.json').read().strip().splitlines()[-1]); print('default', round(d['value']), round(d['e2e']['value']), 'FULL', round(d['full_cycle']['value']), round(d['full_cycle']['e2e']['value']), d['full_cycle']['roofline']['in_situ']['tensor_frac_tf32_sustained'], d.get('parity'))
PY
