#!/bin/bash
# One-GPU validation, as run through `gpurun -- bash tools/gpu_validate.sh`: the GPU test suite, smoke(), the default bench
# line (headline + full cycle + parity + CPU baseline) and the reference arm.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
fc = d['full_cycle']
print('headline', round(d['value']), round(d['e2e']['value']), '| full cycle', round(fc['value']), round(fc['e2e']['value']),
      fc['roofline']['in_situ']['tensor_frac_tf32_sustained'], '| cpu', d.get('cpu_baseline', {}).get('value'))
r = json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('cpu_baseline', {}).get('cores'))
PY
