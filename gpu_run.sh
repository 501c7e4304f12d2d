# round 2, call E (1 GPU): layer-parallel forward in the wide mapping + Actor forward rows
cd /root/repo
O=gpurun_out/r2e; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run t_par python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c3_actor or c4_large_layer_small or c5_unit or odd or actor_more or parameter_changes or learning_disabled"
T=300 run t_small env OGN_SMALL_KERNEL=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit"
T=300 run t_wide env OGN_KERNELS=wide python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit"
T=200 run b_c1 python bench.py --no-cpu --workload c1
T=200 run b_c2 python bench.py --no-cpu --workload c2
T=200 run b_c3 python bench.py --no-cpu --workload c3
T=200 run b_c1_small env OGN_SMALL_KERNEL=1 python bench.py --no-cpu --workload c1
T=200 run b_c2_small env OGN_SMALL_KERNEL=1 python bench.py --no-cpu --workload c2
tail -n 3 $O/t_*.log
grep -h '"metric"' $O/b_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['roofline']['kernels']
    print(d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {a: (b['launches'], round(b['ms_per_launch'],4)) for a, b in k.items()}, d['state_hash']['crc32'])
"
