# helper used during development: the last full single-GPU validation (tests, smoke, default bench, reference arm)
cd /root/repo
mkdir -p gpurun_out/final6
O=gpurun_out/final6
( time timeout 2000 python -m pytest tests -m gpu -x -q ) > $O/tests_gpu.log 2>&1; tail -6 $O/tests_gpu.log | head -3
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench_c4_n1.log 2>&1; grep -a '^{' $O/bench_c4_n1.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms/step %.4f e2e %.1f step_frac %.3f frac %.3f cpu %s launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac'],d['roofline']['frac'],(d.get('cpu_baseline') or {}).get('value'),d['gpu_launches']))"
tail -4 $O/bench_c4_n1.log | grep real
timeout 900 python bench.py --workload c5 --no-cpu > $O/bench_c5.log 2>&1; grep -a '^{' $O/bench_c5.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c5 value %.1f ms/step %.4f step_frac %.3f'%(d['value'],d['ms_per_step'],d['roofline']['step_frac']), {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})"
