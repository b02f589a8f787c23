mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_5_pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_5_bench_default.json 2> gpurun_out/r02_5_bench_default.err
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_5_layers_c3_b256.json > gpurun_out/r02_5_layers_c3_b256.txt 2>&1
tail -n 5 gpurun_out/r02_5_pytest_all.log
grep -E "head|eager" gpurun_out/r02_5_layers_c3_b256.txt | cut -c1-200
bash tools/gpu_run5_sweep.sh
