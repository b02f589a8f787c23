mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread -rf 2>&1 | tail -30 > gpurun_out/r02_10_pytest_all.log
tail -n 4 gpurun_out/r02_10_pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_10_bench_default.json 2> gpurun_out/r02_10_bench_default.err
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/r02_10_layers_c3_b256.json > gpurun_out/r02_10_layers_c3_b256.txt 2>&1
head -14 gpurun_out/r02_10_layers_c3_b256.txt | cut -c1-160
SMK_CONV3_WIN_RING=0 timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 2>&1 | grep -E "eager|N64" | cut -c1-160
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_10_bench_default.json').read().strip().splitlines()[-1]); print('default', round(d['value']), round(d['e2e']['value']), 'FULL', round(d['full_cycle']['value']), round(d['full_cycle']['e2e']['value']), d['full_cycle']['roofline']['in_situ']['tensor_frac_tf32_sustained'])
PY
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_dram_c3_b256_tf32x3.csv python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 1 > gpurun_out/r02_ncu_c3.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_dram_c2_b32_tf32x3.csv python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 1 > gpurun_out/r02_ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_win_kernel -s 2 -c 2 -o gpurun_out/r02_ncu_full_c3_win_H224_K32_N32 python tools/bench_win.py --batch 256 > gpurun_out/r02_ncu_full_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
