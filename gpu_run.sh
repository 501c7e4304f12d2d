cd /root/repo
O=gpurun_out/r2j; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run e2e20 env OGN_BENCH_E2E_TRACE=1 python bench.py --no-cpu --steps 20 --warmup 5
T=300 run e2e60 env OGN_BENCH_E2E_TRACE=1 python bench.py --no-cpu --steps 60 --warmup 5
grep -h "e2e trace" $O/*.log
grep -h '"metric"' $O/e*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['steps'])
"
# small configs with the side stream inside captured graphs
for w in c1 c2 c5; do
  T=300 run b_${w}_side env OGN_SIDE_STREAM=1 python bench.py --no-cpu --workload $w
done
T=300 run b_c3_side env OGN_SIDE_STREAM=1 python bench.py --no-cpu --workload c3
T=300 run t_side env OGN_SIDE_STREAM=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit or parameter_changes or learning_disabled or odd"
tail -n 3 $O/t_side.log
grep -h '"metric"' $O/b_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['roofline']['kernels']
    print(d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {a: (b['launches'], round(b['ms_per_launch'],4)) for a, b in k.items()}, d['state_hash']['crc32'])
"
