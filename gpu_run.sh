cd /root/repo
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d steps/s %.1f ms/step %.3f e2e %.1f step_frac %.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac']), ' '.join('%s %.3fms %.0fGB/s'%(k,v['ms_per_launch'],v['achieved_gbs']) for k,v in r['kernels'].items()))"; }
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 300 --warmup 200 --no-cpu > gpurun_out/n$N.log 2>&1; tail -2 gpurun_out/n$N.log | cut -c1-200; cat gpurun_out/n$N.log | show
done
