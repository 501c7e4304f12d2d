# scratch driver of the last single-GPU measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3i; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=900 run t_par python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c3_actor or c4_large_layer_small or c5_unit or odd or oracle_size_128 or c2_video_full_size"
B="python bench.py --no-cpu --steps 60 --warmup 5"
T=300 run own1_a $B
cp ogmaneo2_b200/libogn_b200.so /tmp/keep.so; cp ogmaneo2_b200/_ab/libogn_b200_allrows.so ogmaneo2_b200/libogn_b200.so
T=300 run own0_a $B
cp /tmp/keep.so ogmaneo2_b200/libogn_b200.so
T=300 run own1_b $B
T=300 run own1_c5 python bench.py --no-cpu --workload c5 --steps 100
cp ogmaneo2_b200/_ab/libogn_b200_allrows.so ogmaneo2_b200/libogn_b200.so
T=300 run own0_c5 python bench.py --no-cpu --workload c5 --steps 100
cp /tmp/keep.so ogmaneo2_b200/libogn_b200.so
tail -n 3 $O/t_par.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3i/own*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), {n:round(v['ms_per_launch'],4) for n,v in k.items()}, d['state_hash']['crc32'])
    except Exception as e: print(f, 'ERR', e)
PY
