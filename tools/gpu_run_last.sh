# scratch driver of the last single-GPU measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3k; mkdir -p $O
timeout 200 env OGN_SMALL_KERNEL=1 python tools/trace_small_step.py --workload c1 > $O/trace_c1.json 2> $O/trace_c1.err; echo "rc=$?"
tail -c 300 $O/trace_c1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3k/trace_c1.json").read().strip().splitlines()[-1])
for s in d[:8]:
    print(s['step'], s['phases'], s['total_us'], [(r['work_us'], r['barrier_us']) for r in s['rows']])
PY
