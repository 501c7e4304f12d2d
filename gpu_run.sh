cd /root/repo
O=gpurun_out/r3c; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=900 run t_small python -m pytest -m gpu -q "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[small_kernel=1]" "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[small_kernel=1-sc_update=2]"
for w in c1 c2; do
T=300 run ${w}_graph python bench.py --no-cpu --workload $w
T=300 run ${w}_small_spread1 env OGN_SMALL_KERNEL=1 python bench.py --no-cpu --workload $w
T=300 run ${w}_small_spread0 env OGN_SMALL_KERNEL=1 OGN_SMALL_SPREAD=0 python bench.py --no-cpu --workload $w
done
tail -n 3 $O/t_small.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r3c/c*_*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['value'],1), round(d['ms_per_step'],4), d['gpu_launches'], d['state_hash']['crc32'])
    except Exception as e: print(f, 'ERR', e)
PY
