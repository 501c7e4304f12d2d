cd /root/repo
O=gpurun_out/r3e; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=1500 run t_all python -m pytest tests -m gpu -q --durations=6
T=600 run d_driver python bench.py --steps 20 --warmup 5
T=400 run d_ref python bench.py --impl reference --steps 20 --warmup 5
tail -n 4 $O/t_all.log; tail -n 1 $O/smoke.log
python - <<'PY'
import json
for f in ("d_driver","d_ref"):
    try:
        d=json.loads(open(f"gpurun_out/r3e/{f}.log").read().strip().splitlines()[-1])
        print(f, round(d['value'],3), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('state_hash') or {}).get('crc32'), d.get('gpu_launches'))
    except Exception as e: print(f, 'ERR', e)
PY
