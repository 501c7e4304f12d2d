cd /root/repo
O=gpurun_out/r2t; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=600 run t_ienc python -m pytest tests/test_image_encoder.py -m gpu -q
T=300 run ienc python tools/bench_image_encoder.py
T=300 run ienc256 python tools/bench_image_encoder.py --size 256 --steps 100
T=300 run tr_128 python tools/trace_image_encoder.py --size 128
CMD="python tools/bench_image_encoder.py --steps 6 --warmup 3"
$CMD > $O/ncu_ienc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_ienc' -s 4 -c 2 -o $O/prof_ienc $CMD > $O/ncu_ienc.log 2>&1
echo "ncu ienc rc=$?" | tee -a $O/summary.txt
tail -n 2 $O/t_ienc.log; tail -n 1 $O/ienc.log; tail -n 1 $O/ienc256.log; cat $O/tr_128.log
