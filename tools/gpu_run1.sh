mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_1_smi.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "3x" --timeout=240 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_1_pytest_3x.log
timeout 900 python -m pytest tests/test_gpu_masking.py tests/test_gpu_parity.py -q --timeout=400 --timeout-method=thread -rf -s 2>&1 | tail -60 > gpurun_out/r02_1_pytest_parity.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_masking.py 2>&1 | tail -40 > gpurun_out/r02_1_pytest_rest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_1_bench_x3.json 2> gpurun_out/r02_1_bench_x3.err
timeout 600 python bench.py --steps 20 --warmup 5 --precision tf32 --no-full-cycle --no-cpu-baseline > gpurun_out/r02_1_bench_tf32.json 2> gpurun_out/r02_1_bench_tf32.err
tail -5 gpurun_out/r02_1_pytest_3x.log gpurun_out/r02_1_pytest_parity.log gpurun_out/r02_1_pytest_rest.log
tail -c 600 gpurun_out/r02_1_bench_x3.err
