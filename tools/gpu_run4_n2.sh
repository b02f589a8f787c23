# 2-GPU frame-shard bench with the NCCL all-gather inside the timed region (charged 2x)
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_4_bench_n2.json 2> gpurun_out/r02_4_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-affinity --no-full-cycle > gpurun_out/r02_4_bench_n2_noaff.json 2> gpurun_out/r02_4_bench_n2_noaff.err
tail -c 400 gpurun_out/r02_4_bench_n2.err
python - <<'PY'
import json
for f in ('r02_4_bench_n2','r02_4_bench_n2_noaff'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value']),'e2e',round(d['e2e']['value']), 'gather',d.get('gather'), 'no_gather',d.get('no_gather'), d['config'].get('host_affinity'))
        fc=d.get('full_cycle')
        if fc: print('  full', round(fc['value']), round(fc['e2e']['value']), fc.get('gather'), fc.get('no_gather'))
    except Exception as e: print(f,'ERR',e)
PY
