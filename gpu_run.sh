cd /root/repo
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('steps/s %.1f ms/step %.3f e2e %.1f step_frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac']), ' '.join('%s %.3fms %.0fGB/s'%(k,v['ms_per_launch'],v['achieved_gbs']) for k,v in r['kernels'].items()), 'pairs/step', r['sc_pairs_rewritten_per_step'])"; }
echo steady; python bench.py --steps 300 --warmup 200 --no-cpu 2>&1 | show
