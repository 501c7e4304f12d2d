cd /root/repo
mkdir -p gpurun_out
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value %.1f ms/step %.4f e2e %.1f step_frac %.3f cpu %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac'], (d.get('cpu_baseline') or {}).get('value')), ' '.join('%s %.4fms x%d %.0fGB/s'%(k,v['ms_per_launch'],v['launches'],v['achieved_gbs'] or 0) for k,v in r['kernels'].items()))"; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; tail -4 gpurun_out/tests.log
for w in c1 c2 c3; do
timeout 600 python bench.py --workload $w > gpurun_out/w_$w.log 2>&1; echo $w; tail -2 gpurun_out/w_$w.log | cut -c1-300 | grep -v '^{' ; cat gpurun_out/w_$w.log | show
done
timeout 900 python bench.py --workload c5 --ensemble 512 > gpurun_out/w_c5.log 2>&1; echo c5; tail -2 gpurun_out/w_c5.log | cut -c1-300 | grep -v '^{'; cat gpurun_out/w_c5.log | show
timeout 600 python bench.py --steps 100 --warmup 100 --no-cpu > gpurun_out/w_c4.log 2>&1; echo c4; cat gpurun_out/w_c4.log | show
