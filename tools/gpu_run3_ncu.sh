# ncu evidence for profiles/ (1 GPU; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
# (1) launch list + DRAM traffic of one eager pass, headline workload (configs[1], B=32, tf32x3)
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_dram_c2_b32_tf32x3.csv python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 1 > gpurun_out/r02_ncu_c2.log 2>&1
# (2) the same for the full cycle at B=256
timeout 1500 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_ncu_dram_c3_b256_tf32x3.csv python tools/profile_layers.py --batch 256 --generator --precision tf32x3 --steps 1 > gpurun_out/r02_ncu_c3.log 2>&1
# (3) --set full of the generator's dominant conv shapes: 14^2 512->512 (BN=256 tile) and 224^2 32->32
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/r02_ncu_full_c3_H14_K512_N512 python tools/bench_pw.py --gen --batch 256 --reps 2 --only c3_H14_K512_N512 > gpurun_out/r02_ncu_full_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/r02_ncu_full_c3_H224_K32_N32 python tools/bench_pw.py --gen --batch 256 --reps 2 --only c3_H224_K32_N32 > gpurun_out/r02_ncu_full_b.log 2>&1
# (4) --set full of the encoder's two tensor-core kernels in the 3xTF32 path, inside one eager pass
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xdw_kernel|gemm_tc_kernel" -s 60 -c 12 -o gpurun_out/r02_ncu_full_encoder_x3 python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 1 > gpurun_out/r02_ncu_full_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
