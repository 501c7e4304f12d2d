cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
N=${NGPU:-8}
nproc > $O/nproc.txt; lscpu | head -20 >> $O/nproc.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 150 --warmup 50 --burnin 0 > $O/bench_n${N}c.log 2>&1
grep -a '^{' $O/bench_n${N}c.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d steps/s %.1f ms/step %.3f e2e %.1f step_frac %.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac']))
for i,x in enumerate(d.get('ranks') or []): print('   rank',i,'host_enqueue %.3f wait_us %.1f kernel_ms %s'%(x['host_enqueue_ms_per_step'],x['exchange_wait_us_per_step'],x['kernel_ms']))"
cat $O/nproc.txt | head -12
