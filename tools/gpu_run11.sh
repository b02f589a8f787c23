mkdir -p gpurun_out
for ws in 0 1; do
  SMK_XDW_WORKER_SPLIT=$ws timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_11_ws$ws.json 2> gpurun_out/r02_11_ws$ws.err
  SMK_XDW_WORKER_SPLIT=$ws timeout 600 python tools/profile_layers.py --batch 32 --precision tf32x3 --steps 3 > gpurun_out/r02_11_layers_c2_ws$ws.txt 2>&1
  SMK_XDW_WORKER_SPLIT=$ws timeout 600 python tools/profile_layers.py --batch 256 --precision tf32x3 --steps 2 > gpurun_out/r02_11_layers_c2b256_ws$ws.txt 2>&1
done
SMK_XDW_WORKER_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_11_ws0b.json 2>/dev/null
SMK_XDW_WORKER_SPLIT=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_11_ws1b.json 2>/dev/null
python - <<'PY'
import json
for n in ('ws0','ws1','ws0b','ws1b'):
    try:
        d=json.loads(open('gpurun_out/r02_11_%s.json'%n).read().strip().splitlines()[-1]); print(n, round(d['value']), round(d['e2e']['value']), d.get('spread'))
    except Exception as e: print(n, 'ERR', e)
PY
for ws in 0 1; do echo ws$ws; grep -E "eager|xdw" gpurun_out/r02_11_layers_c2_ws$ws.txt | head -8 | cut -c1-150; grep -E "eager|xdw" gpurun_out/r02_11_layers_c2b256_ws$ws.txt | head -6 | cut -c1-150; done
