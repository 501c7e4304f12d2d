cd /root/repo
O=gpurun_out/r3b; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=600 run t_ienc python -m pytest tests/test_image_encoder.py -m gpu -q
T=300 run pipe256 python tools/bench_image_pipeline.py --size 256
T=300 run pipe128 python tools/bench_image_pipeline.py --size 128
T=300 run side1 env OGN_SIDE_STREAM=1 python bench.py --no-cpu --steps 60 --warmup 5
T=300 run side0 python bench.py --no-cpu --steps 60 --warmup 5
tail -n 5 $O/t_ienc.log; tail -n 1 $O/pipe256.log; tail -n 1 $O/pipe128.log
python - <<'PY'
import json
for f in ("side1","side0"):
    try:
        d=json.loads(open(f"gpurun_out/r3b/{f}.log").read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d['state_hash']['crc32'])
    except Exception as e: print(f,'ERR',e)
PY
