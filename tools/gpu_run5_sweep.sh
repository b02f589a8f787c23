mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-full-cycle --no-cpu-baseline --no-parity"
run() { name=$1; shift; env "$@" timeout 600 $B > gpurun_out/r02_5_$name.json 2> gpurun_out/r02_5_$name.err; }
run base SMK_NOP=1
run xdwslots74 SMK_XDW_SLOTS=74
run xdwslots110 SMK_XDW_SLOTS=110
run minres15 SMK_XDW_MIN_RES=15
run minres29 SMK_XDW_MIN_RES=29
run nopair SMK_ENC_PAIR=0
run pdl SMK_PDL=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_5_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('r02_5_')[1], round(d['value']), round(d['e2e']['value']), d['config']['execution'][:34])
    except Exception as e: print(f,'ERR',e)
PY
for c in 1 2 3; do SMK_WIN_CTAS=$c timeout 300 python tools/bench_win.py 2>&1 | tail -3; done
SMK_WIN_CTAS=2 timeout 300 python tools/bench_win.py --cin 64 2>&1 | tail -3
