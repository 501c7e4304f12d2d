cd /root/repo
O=gpurun_out/r2k; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=600 run t_par python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not variant and not random and not full_size and not large_layer_1 and not 224"
T=600 run t_feat python -m pytest tests/test_gpu_features.py tests/test_image_encoder.py tests/test_dropin_cpp.py -m gpu -q -x -k "not full_size"
for w in c1 c2 c3 c5; do T=300 run b_$w python bench.py --no-cpu --workload $w; done
T=300 run b_c4 python bench.py --no-cpu --steps 60 --warmup 5
T=600 run d_default python bench.py --steps 20 --warmup 5
tail -n 3 $O/t_*.log
grep -h '"metric"' $O/b_*.log $O/d_default.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['roofline']['kernels']
    print(d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['step_frac'],3), {a: (b['launches'], round(b['ms_per_launch'],4)) for a, b in k.items()}, d['state_hash']['crc32'])
"
