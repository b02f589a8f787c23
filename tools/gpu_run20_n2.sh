mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 tools/check_gather.py 2>&1 | grep -E "gather backend|Error|error|Traceback" | tail -6
for be in p2p; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 5 --gather-backend $be --no-cpu-baseline --no-parity > gpurun_out/r02_20_bench_n2_$be.json 2> gpurun_out/r02_20_bench_n2_$be.err
tail -c 300 gpurun_out/r02_20_bench_n2_$be.err | grep -v "^\*\|OMP" 
done
python - <<'PY'
import json
for f in ('r02_20_bench_n2_p2p','r02_20_bench_n2_nccl'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),'e2e',round(d['e2e']['value']), 'gather',d.get('gather',{}).get('backend'), 'no_gather',round(d['no_gather']['value']), round(d['no_gather']['e2e_value']))
        fc=d.get('full_cycle')
        if fc: print('  full', round(fc['value']), round(fc['e2e']['value']), 'no_gather', round(fc['no_gather']['value']), round(fc['no_gather']['e2e_value']))
    except Exception as e: print(f,'ERR',e)
PY
