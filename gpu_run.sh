cd /root/repo
mkdir -p gpurun_out/actor
O=gpurun_out/actor
timeout 1200 python -m pytest tests -m gpu -x -q -k "actor or c3 or golden or parameter_changes" > $O/tests2.log 2>&1; tail -4 $O/tests2.log
show() { grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s %.1f ms/step %.4f e2e %.1f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['gpu_launches']), {k:round(v['ms_per_launch'],4) for k,v in d['roofline']['kernels'].items()})"; }
timeout 600 python bench.py --workload c3 --no-cpu 2>&1 | tee $O/c3b.log | show
