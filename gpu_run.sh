cd /root/repo
O=gpurun_out/r3a; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
for v in vec1 vec2 wide wide2 stage tma; do
T=300 run c5_$v env OGN_LARGE_VARIANT=$v python bench.py --no-cpu --workload c5 --steps 100
done
T=300 run c5_stage1 env OGN_LARGE_VARIANT=stage OGN_STAGE_CHUNKS=1 python bench.py --no-cpu --workload c5 --steps 100
T=300 run c5_stage2 env OGN_LARGE_VARIANT=stage OGN_STAGE_CHUNKS=2 python bench.py --no-cpu --workload c5 --steps 100
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3a/c5_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), {n:round(v['ms_per_launch'],4) for n,v in k.items()}, d['state_hash']['crc32'])
    except Exception as e: print(f, 'ERR', e)
PY
