cd /root/repo
mkdir -p gpurun_out
show() { tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d steps/s %.1f ms/step %.3f e2e %.1f step_frac %.3f upd %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],r['step_frac'],r['sc_pairs_rewritten_per_step']), ' '.join('%s %.3fms %.0fGB/s'%(k,v['ms_per_launch'],v['achieved_gbs'] or 0) for k,v in r['kernels'].items()))"; }
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "peer or sharded" > gpurun_out/tests_peer.log 2>&1; tail -15 gpurun_out/tests_peer.log
N=${NGPU:-2}
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 100 --no-cpu --exchange $ex > gpurun_out/n${N}_$ex.log 2>&1; tail -3 gpurun_out/n${N}_$ex.log | cut -c1-300; echo $ex; cat gpurun_out/n${N}_$ex.log | show
done
