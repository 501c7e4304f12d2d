# round 2, final 8-GPU call (short): C4 strong scaling and C5 weak scaling at the default step counts
cd /root/repo
O=gpurun_out/r2l; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
T=200 run c4_n8 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8
T=200 run c4_n8_s20 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5
T=200 run c4_n4_s20 $TR --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5
T=200 run c4_n2_s20 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5
T=300 run c5_n8 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --workload c5
grep -h '"metric"' $O/c*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:10], d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['state_hash']['crc32'], d['steps'], d['warmup'])
"
