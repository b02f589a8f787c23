mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/bench_peer.py 2>&1 | grep -E "idle|loaded|Error|error|Traceback" | tee gpurun_out/r02_18_peer_n2.txt
nvidia-smi topo -m 2>/dev/null | head -6
