#!/bin/bash
# One-GPU profiling pass (`gpurun -- bash tools/gpu_profile.sh`): per-layer device times from the built-in event profiler,
# the ncu launch lists with DRAM bytes for configs[1] / configs[2] (-> tools/ncu_traffic.py -> profiles/*_dram_traffic_*.json)
# and `--set full` captures of the dominant kernels (-> tools/ncu_summary.py).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
timeout 600 python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 3 --out gpurun_out/layers_c2_b32.json > gpurun_out/layers_c2_b32.txt 2>&1
timeout 600 python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 2 --out gpurun_out/layers_c3_b256.json > gpurun_out/layers_c3_b256.txt 2>&1
head -14 gpurun_out/layers_c3_b256.txt | cut -c1-160
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_dram_c3_b256_tf32x3.csv python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 1 > gpurun_out/ncu_c3.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/ncu_dram_c2_b32_tf32x3.csv python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 1 > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3_win_kernel -s 2 -c 2 -o gpurun_out/ncu_full_c3_win_H224_K32_N32 python tools/bench_win.py --batch 256 > gpurun_out/ncu_full_win.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xdw_kernel -s 3 -c 1 -o gpurun_out/ncu_full_xdw3x_H112_K16_N64_s2 python tools/bench_xdw.py --x3 --batch 256 --reps 2 --only H112_K16 > gpurun_out/ncu_full_xdw.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
