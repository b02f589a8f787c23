mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > gpurun_out/r02_22_bench_n8.json 2> gpurun_out/r02_22_bench_n8.err
tail -c 300 gpurun_out/r02_22_bench_n8.err | grep -v "^\*\|OMP"
python - <<'PY'
import json
for f in ('r02_22_bench_n8',):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),'e2e',round(d['e2e']['value']), 'gather',d.get('gather',{}).get('backend'), d.get('gather',{}).get('recv_gbs_per_rank'), 'no_gather',round(d['no_gather']['value']), round(d['no_gather']['e2e_value']))
        fc=d.get('full_cycle')
        if fc: print('  full', round(fc['value']), round(fc['e2e']['value']), 'no_gather', round(fc['no_gather']['value']), round(fc['no_gather']['e2e_value']))
    except Exception as e: print(f,'ERR',e)
PY
