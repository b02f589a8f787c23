mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/check_gather.py 2>&1 | grep -E "gather backend|Error|error|Traceback" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_17_bench_n8.json 2> gpurun_out/r02_17_bench_n8.err
tail -c 300 gpurun_out/r02_17_bench_n8.err | grep -v "^\*\|OMP"
python - <<'PY'
import json
for f in ('r02_17_bench_n8',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),'e2e',round(d['e2e']['value']), 'gather',d.get('gather'), 'no_gather',d.get('no_gather'), d.get('clocks'))
        fc=d.get('full_cycle')
        if fc: print('  full', round(fc['value']), round(fc['e2e']['value']), fc.get('gather'), fc.get('no_gather'))
    except Exception as e: print(f,'ERR',e)
PY
