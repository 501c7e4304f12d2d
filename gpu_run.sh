cd /root/repo
mkdir -p gpurun_out
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d steps/s %.1f ms/step %.3f e2e %.1f step_frac %.3f upd %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac'],r['sc_pairs_rewritten_per_step']), ' '.join('%s %.3fms %.0fGB/s'%(k,v['ms_per_launch'],v['achieved_gbs'] or 0) for k,v in r['kernels'].items()))"; }
run() { name=$1; shift; env "$@" python bench.py --steps 100 --warmup 100 --no-cpu > gpurun_out/bench_$name.log 2>&1; echo $name "$@"; cat gpurun_out/bench_$name.log | show; }
python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; tail -5 gpurun_out/tests.log
run u1 OGN_UPD_MODE=1
run u0 OGN_UPD_MODE=0
run u2 OGN_UPD_MODE=2
