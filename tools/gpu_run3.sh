mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_masking.py "tests/test_gpu_parity.py::test_benched_pipeline_parity_configs1" tests/test_gpu_parity.py::test_generator_tf32_tensor_core_path -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -40 > gpurun_out/r02_3_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r02_3_breakdown.json > gpurun_out/r02_3_bench_default.json 2> gpurun_out/r02_3_bench_default.err
SMK_X3_DEEP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_3_bench_nodeep.json 2> gpurun_out/r02_3_bench_nodeep.err
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_3_layers_c3_b256.json > gpurun_out/r02_3_layers_c3_b256.txt 2>&1
timeout 600 python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 5 --out gpurun_out/r02_3_layers_c2_b32.json > gpurun_out/r02_3_layers_c2_b32.txt 2>&1
tail -n 4 gpurun_out/r02_3_pytest.log
python - <<'PY'
import json
for f in ('r02_3_bench_default','r02_3_bench_nodeep'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, round(d['value']), round(d['e2e']['value']), d.get('parity') and d['parity'].get('params_rel'), d.get('full_cycle') and round(d['full_cycle']['value']))
    except Exception as e: print(f,'ERR',e)
PY
bash tools/gpu_run3_ncu.sh
