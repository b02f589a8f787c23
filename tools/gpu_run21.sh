mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -5 > gpurun_out/r02_21_pytest_all.log
tail -n 3 gpurun_out/r02_21_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_21_bench_default.json 2> gpurun_out/r02_21_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_21_bench_reference.json 2> gpurun_out/r02_21_bench_reference.err
tail -c 600 gpurun_out/r02_21_bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_21_bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['value']), round(d['e2e']['value']), 'FULL', round(d['full_cycle']['value']), round(d['full_cycle']['e2e']['value']), d['full_cycle']['roofline']['in_situ']['tensor_frac_tf32_sustained'], 'cpu', d.get('cpu_baseline'))
PY
