mkdir -p gpurun_out
for c in 1 2 3; do SMK_WIN_CTAS=$c timeout 300 python tools/bench_win.py 2>&1 | tail -3; done
SMK_WIN_CTAS=2 timeout 300 python tools/bench_win.py --cin 64 2>&1 | tail -3
for s in 50 74; do SMK_XDW_SLOTS=$s timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_6_slots$s.json 2>/dev/null; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity > gpurun_out/r02_6_default.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_6_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('r02_6_')[1], round(d['value']), round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
