mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/bench_peer_load.py 2>&1 | grep -E "push|copies|Error|error|Traceback" | tee gpurun_out/r02_19_peer_load_n2.txt
