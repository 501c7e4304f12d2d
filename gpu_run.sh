cd /root/repo
O=gpurun_out/r3d; mkdir -p $O
timeout 300 env OGN_SMALL_KERNEL=1 python tools/trace_small_step.py --workload c2 > $O/trace_c2.json 2> $O/trace_c2.err; echo "rc=$?"
tail -c 600 $O/trace_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3d/trace_c2.json").read().strip().splitlines()[-1])
for s in d:
    print(s['step'], s['phases'], s['total_us'])
    for r in s['rows']: print('   ', r)
PY
