cd /root/repo
mkdir -p gpurun_out/final3
O=gpurun_out/final3
N=2
show() { grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d value %.1f ms/step %.4f e2e %.1f step_frac %.3f scaling %s | %s'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac'],d['scaling'],d['config']['parallelism']))"; }
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
run --steps 100 --warmup 50 --no-cpu 2>&1 | tee $O/n2_peer.log | show
run --steps 100 --warmup 50 --no-cpu --exchange nccl 2>&1 | tee $O/n2_nccl.log | show
run --workload c5 --ensemble 128 --no-cpu 2>&1 | tee $O/n2_c5.log | show
python bench.py --workload c5 --ensemble 128 --no-cpu 2>&1 | tee $O/n1_c5.log | show
tail -3 $O/n2_c5.log | cut -c1-300
