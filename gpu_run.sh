cd /root/repo
O=gpurun_out/r3g; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
B="python bench.py --no-cpu --size 184 --steps 300"
T=200 run s184_default $B
T=200 run s184_t128 env OGN_VEC_THREADS=128 $B
T=200 run s184_t64 env OGN_VEC_THREADS=64 $B
T=200 run s184_stage env OGN_LARGE_VARIANT=stage $B
T=200 run s184_stage_t128 env OGN_LARGE_VARIANT=stage OGN_VEC_THREADS=128 $B
T=200 run s184_side env OGN_SIDE_STREAM=1 $B
T=200 run s184_side_t128 env OGN_SIDE_STREAM=1 OGN_VEC_THREADS=128 $B
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3g/s184_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), {n:round(v['ms_per_launch'],4) for n,v in k.items()}, d['state_hash']['crc32'])
    except Exception as e: print(f, 'ERR', e)
PY
