# round 2, call H (1 GPU): update kernel with 64-byte full stores, then the whole GPU suite
cd /root/repo
O=gpurun_out/r2h; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run t_upd3 env OGN_SC_UPDATE=3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_video_small or c4_large_layer_small or c5_unit or c1_sine"
T=300 run t_upd2 env OGN_SC_UPDATE=2 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_video_small or c4_large_layer_small or c5_unit"
T=300 run u_v1 python bench.py --no-cpu --steps 60 --warmup 5
T=300 run u_v3 env OGN_SC_UPDATE=3 python bench.py --no-cpu --steps 60 --warmup 5
T=300 run u_v3b1 env OGN_SC_UPDATE=3 OGN_UPD2_BATCHES=1 python bench.py --no-cpu --steps 60 --warmup 5
T=300 run u_v2 env OGN_SC_UPDATE=2 python bench.py --no-cpu --steps 60 --warmup 5
T=1500 run t_all python -m pytest tests -m gpu -q
tail -n 3 $O/t_*.log
grep -h '"metric"' $O/u_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['roofline']['kernels']
    print(d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {a: (b['launches'], round(b['ms_per_launch'],4)) for a, b in k.items()}, d['state_hash']['crc32'])
"
