cd /root/repo
mkdir -p gpurun_out/final4
O=gpurun_out/final4
( time timeout 2000 python -m pytest tests -m gpu -x -q ) > $O/tests_gpu.log 2>&1; tail -6 $O/tests_gpu.log | head -3
show() { grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms/step %.4f e2e %.1f step_frac %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac']), {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})"; }
for w in c1 c2 c3; do echo $w; timeout 600 python bench.py --workload $w 2>&1 | tee $O/bench_$w.log | show; done
echo c5; timeout 900 python bench.py --workload c5 2>&1 | tee $O/bench_c5.log | show
echo c4; timeout 900 python bench.py --no-cpu 2>&1 | tee $O/bench_c4.log | show
