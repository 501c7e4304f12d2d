cd /root/repo
mkdir -p gpurun_out/final5
O=gpurun_out/final5
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "peer or sharded" > $O/shard.log 2>&1; tail -5 $O/shard.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-cpu 2>&1 | grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=%d value %.1f ms/step %.4f e2e %.1f h2d %d d2h %d'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['h2d_bytes_per_step'],d['e2e']['d2h_bytes_per_step']))"
