# scratch driver of the last measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3r; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "c1_contract_length" > $O/t_c1.log 2>&1; echo "t_c1 rc=$?"
timeout 200 python bench.py --no-cpu --workload c1 > $O/b_c1.log 2>&1; echo "b_c1 rc=$?"
tail -n 3 $O/t_c1.log | cut -c1-250
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r3r/b_c1.log").read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], d.get('c1_last1000_prediction_errors'), d['state_hash']['crc32'])
PY
