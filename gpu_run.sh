cd /root/repo
O=gpurun_out/r2m; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=300 run ienc python tools/bench_image_encoder.py
T=1500 run t_all python -m pytest tests -m gpu -q
CMD="env OGN_LARGE_VARIANT=tma python bench.py --no-cpu --steps 10 --warmup 5 --burnin 100"
$CMD > $O/ncu_tma_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_(sc_forward|pred_step)_t' -s 200 -c 2 -o $O/prof_c4_512_tma $CMD > $O/ncu_tma.log 2>&1
echo "ncu tma rc=$?" | tee -a $O/summary.txt
tail -n 3 $O/t_all.log; cat $O/ienc.log | tail -2
