mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_14_topo.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/check_gather.py 2>&1 | grep -E "gather backend|Error|error" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_14_bench_n8.json 2> gpurun_out/r02_14_bench_n8.err
tail -c 400 gpurun_out/r02_14_bench_n8.err
python - <<'PY'
import json
for f in ('r02_14_bench_n8',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),'e2e',round(d['e2e']['value']), 'gather',d.get('gather'), 'no_gather',d.get('no_gather'), d.get('clocks'))
        fc=d.get('full_cycle')
        if fc: print('  full', round(fc['value']), round(fc['e2e']['value']), fc.get('gather'), fc.get('no_gather'))
    except Exception as e: print(f,'ERR',e)
PY
