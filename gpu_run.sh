cd /root/repo
O=gpurun_out/r2u; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run smoke python -c "import __graft_entry__ as g; g.smoke()"
T=1500 run t_all python -m pytest tests -m gpu -q --durations=6
T=600 run d_default python bench.py
T=400 run d_ref python bench.py --impl reference
T=300 run ienc python tools/bench_image_encoder.py
T=300 run ienc256 python tools/bench_image_encoder.py --size 256 --steps 100
T=1500 run san_mem compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest -q -x tests/test_image_encoder.py tests/test_gpu_parity.py -m gpu -k "(test_device_matches_checker and not bench_size) or test_device_ties or test_c1_sine or test_c5_unit or test_c3_actor or shape1 or kernels=tma or kernels=wide2 or (small_kernel=1 and not sc_update)"
tail -n 4 $O/t_all.log; tail -n 6 $O/san_mem.log
python - <<'PY'
import json
for f in ("d_default","d_ref","ienc","ienc256"):
    try:
        d=json.loads(open(f"gpurun_out/r2u/{f}.log").read().strip().splitlines()[-1])
        print(f, round(d['value'],3), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'), (d.get('state_hash') or {}).get('crc32'))
    except Exception as e: print(f, 'ERR', e)
PY
