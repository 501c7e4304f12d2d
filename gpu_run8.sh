cd /root/repo
O=gpurun_out/r2g; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 env OGN_TIMELINE=24 $TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 64 --warmup 5 > $O/c4_n8_timeline.log 2>&1; echo rc=$?
grep "ogn timeline\] dev 3" $O/c4_n8_timeline.log | head -30
grep -h '"metric"' $O/c4_n8_timeline.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1))
"
