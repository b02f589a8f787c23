mkdir -p gpurun_out
timeout 300 python tools/bench_xdw.py --x3 --batch 256 --reps 5 > gpurun_out/r02_12_xdw_b256_x3.txt 2>&1
timeout 300 python tools/bench_xdw.py --batch 256 --reps 5 > gpurun_out/r02_12_xdw_b256_tf32.txt 2>&1
cat gpurun_out/r02_12_xdw_b256_x3.txt | head -8; head -4 gpurun_out/r02_12_xdw_b256_tf32.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xdw_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_full_xdw3x_H112_K16_N64_s2 python tools/bench_xdw.py --x3 --batch 256 --reps 2 --only H112_K16 > gpurun_out/r02_12_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xdw_kernel -s 3 -c 1 -o gpurun_out/r02_ncu_full_xdw3x_H56_K24_N72_s1 python tools/bench_xdw.py --x3 --batch 256 --reps 2 --only H56_K24_N72_s1 > gpurun_out/r02_12_ncu_b.log 2>&1
ls -la gpurun_out/*xdw3x*.ncu-rep
