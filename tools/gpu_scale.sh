#!/bin/bash
# Multi-GPU run (`gpurun --gpus N -- bash tools/gpu_scale.sh N`): both gather backends against a plain NCCL all_gather
# (tools/check_gather.py), the copy-engine push bandwidth (tools/bench_peer.py), then the bench line at N GPUs.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29541 tools/check_gather.py 2>&1 | grep -E "gather backend|Error|error|Traceback" | tail -6
timeout 300 $RUN --master-port 29542 tools/bench_peer.py 2>&1 | grep -E "idle|loaded|Error|error|Traceback" | tee gpurun_out/peer_bandwidth_n$N.txt
timeout 600 $RUN --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_n%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
g, ng, fc = d.get('gather', {}), d.get('no_gather', {}), d.get('full_cycle', {})
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), g.get('backend'), 'no_gather', round(ng.get('value', 0)), round(ng.get('e2e_value', 0)))
if fc:
    print('full', round(fc['value']), round(fc['e2e']['value']), 'no_gather', round(fc['no_gather']['value']), round(fc['no_gather']['e2e_value']))
PY
