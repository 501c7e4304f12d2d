# round 2, call A: new kernels (single-launch step, wide, tma, update v2) — correctness first, then A/B timings
cd /root/repo
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
free -g > $O/host.txt; nproc >> $O/host.txt
run() { name=$1; shift; timeout "${T:-600}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
# 1. quick sanity of each new piece (separate processes: a device fault must not poison the others)
T=300 run t_small python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit"
T=300 run t_nosmall env OGN_SMALL_KERNEL=0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit"
T=300 run t_tma env OGN_KERNELS=tma python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit"
T=300 run t_upd1 env OGN_KERNELS=vec1 OGN_SC_UPDATE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_video_small or c4_large_layer_small"
# 2. the whole GPU suite
T=1500 run t_all python -m pytest tests -m gpu -q
# 3. C4 full size A/B (no CPU leg): baseline vec1 + update v1, update v2, tma
T=400 run b_c4_upd1 env OGN_SC_UPDATE=1 python bench.py --no-cpu --steps 100 --warmup 5
T=400 run b_c4_upd2 python bench.py --no-cpu --steps 100 --warmup 5
T=400 run b_c4_tma env OGN_LARGE_VARIANT=tma python bench.py --no-cpu --steps 100 --warmup 5
# 4. small configs: single-launch step vs per-phase launches
for w in c1 c2 c5; do
  T=400 run b_${w}_small python bench.py --no-cpu --workload $w
  T=400 run b_${w}_launch env OGN_SMALL_KERNEL=0 python bench.py --no-cpu --workload $w
  T=400 run b_${w}_launch_vec2 env OGN_SMALL_KERNEL=0 OGN_SMALL_VARIANT=vec2 python bench.py --no-cpu --workload $w
done
tail -n 3 $O/t_*.log
grep -h '"metric"' $O/b_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); k = d['roofline']['kernels']
    print(d['config']['workload'][:12], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), {a: round(b['ms_per_launch'],4) for a, b in k.items()}, d['state_hash']['crc32'])
"
