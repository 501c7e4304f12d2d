# round 2, call B (1 GPU): new tests, update-kernel sweep, 1/8-size variants, stress streams, CPU baselines, ncu evidence
cd /root/repo
O=gpurun_out/r2b; mkdir -p $O
run() { name=$1; shift; timeout "${T:-600}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=400 run t_new python -m pytest tests/test_gpu_features.py -m gpu -q -k "readback or with_actor or get_set_state or deterministic or checkpoint"
T=300 run t_par python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c1_sine or c2_video_small or c4_large_layer_small or c5_unit or c3_actor"
T=300 run t_par_launch env OGN_SMALL_KERNEL=0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "c2_video_small or c4_large_layer_small or c5_unit"
# update kernel sweep, C4 full size and C5 (per-phase launches)
T=300 run u_c4_v1 env OGN_SC_UPDATE=1 python bench.py --no-cpu --steps 60 --warmup 5
for b in 1 2 4 8; do T=300 run u_c4_v2b$b env OGN_UPD2_BATCHES=$b python bench.py --no-cpu --steps 60 --warmup 5; done
T=400 run u_c5_v1 env OGN_SMALL_KERNEL=0 OGN_SC_UPDATE=1 python bench.py --no-cpu --workload c5 --steps 100
for b in 2 4 8; do T=400 run u_c5_v2b$b env OGN_SMALL_KERNEL=0 OGN_UPD2_BATCHES=$b python bench.py --no-cpu --workload c5 --steps 100; done
# 1/8-size launches (what one rank of 8 runs): octet-per-pack vs warp-per-pack
T=300 run s184_vec1 python bench.py --no-cpu --size 184 --steps 300
T=300 run s184_wide env OGN_SMALL_PACKS=60000 python bench.py --no-cpu --size 184 --steps 300
T=300 run s184_wide2 env OGN_SMALL_PACKS=60000 OGN_SMALL_VARIANT=wide2 python bench.py --no-cpu --size 184 --steps 300
T=300 run s184_t128 env OGN_VEC_THREADS=128 python bench.py --no-cpu --size 184 --steps 300
# stress streams at full size
T=400 run st_iid python bench.py --no-cpu --stream iid
T=400 run st_coherent python bench.py --no-cpu --stream coherent
# C2 with the deeper wide variant
T=300 run c2_wide2 env OGN_SMALL_KERNEL=0 OGN_SMALL_VARIANT=wide2 python bench.py --no-cpu --workload c2
T=300 run c1_wide2 env OGN_SMALL_KERNEL=0 OGN_SMALL_VARIANT=wide2 python bench.py --no-cpu --workload c1
# the driver's own command, CPU leg and GPU-parity-on-sample included
T=600 run d_default python bench.py --steps 20 --warmup 5
# CPU baselines on this box's host cores
T=600 run r_c4_128 python bench.py --impl reference --steps 20 --warmup 3
T=900 run r_c4_224 python bench.py --impl reference --steps 20 --warmup 3 --cpu-size 224
T=600 run r_c5 python bench.py --impl reference --workload c5 --steps 20 --warmup 3
T=600 run r_c2 python bench.py --impl reference --workload c2 --steps 20 --warmup 3
T=300 run r_c1 python bench.py --impl reference --workload c1 --steps 2000 --warmup 100
# ncu: launch list of the timed region, then one full capture of the four step kernels (same command, plain run first)
CMD="python bench.py --no-cpu --steps 10 --warmup 5 --burnin 100"
$CMD > $O/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(sc|pred)' -s 400 -c 48 --csv --log-file $O/launches_c4_512.csv $CMD > $O/ncu_launch.log 2>&1
echo "ncu launches rc=$?" | tee -a $O/summary.txt
ncu --set full --clock-control none --import-source on -k regex:'k_(sc|pred)' -s 400 -c 4 -o $O/prof_c4_512 $CMD > $O/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $O/summary.txt
tail -n 2 $O/t_*.log
grep -h '"metric"' $O/[usc]*.log $O/d_default.log $O/r_*.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d.get('impl') == 'reference':
        print('REF', d['config']['workload'][:60], round(d['value'],4), d['cpu_baseline']['sample'][:90]); continue
    k = d['roofline']['kernels']
    print(d['config']['workload'][:14], d['config']['stream'], round(d['value'],1), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['step_frac'],3), {a: round(b['ms_per_launch'],4) for a, b in k.items()}, d['roofline']['sc_pairs_rewritten_per_step'][:1], d['state_hash']['crc32'], (d.get('cpu_baseline') or {}).get('gpu_parity_on_sample'))
"
