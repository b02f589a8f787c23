SMK_GATHER_PACK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/check_gather.py 2>&1 | grep -E "gather backend|Error|error|Traceback" | tail -6
SMK_GATHER_PACK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --windows 5 --no-cpu-baseline --no-parity --no-full-cycle 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('packed n2', round(d['value']), round(d['e2e']['value']), d['gather']['backend'], round(d['no_gather']['value']), round(d['no_gather']['e2e_value']))"
