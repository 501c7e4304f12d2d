# round 2, call F (8 GPUs, short): C4 strong scaling at the default step counts + two scheduling knobs, C5 at its stated scale
cd /root/repo
O=gpurun_out/r2f; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
T=200 run c4_n8 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8
T=200 run c4_n8_nohoist env OGN_PRED_HOIST=0 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 150 --warmup 5
T=200 run c4_n8_t128 env OGN_VEC_THREADS=128 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --steps 150 --warmup 5
T=300 run c5_n8 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --workload c5
grep -h '"metric"' $O/c*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:10], d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), d['state_hash']['crc32'], d['steps'], d['warmup'])
    for r in (d.get('ranks') or [])[:3]:
        print('   ', round(r['step_ms'],4), r['kernel_ms'], r.get('exchange_wait_us_per_step') and round(r['exchange_wait_us_per_step'],1), round(r['host_enqueue_ms_per_step'],3))
"
