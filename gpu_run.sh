cd /root/repo
mkdir -p gpurun_out/final
O=gpurun_out/final
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt
( time python -m pytest tests -m gpu -x -q ) > $O/tests_gpu.log 2>&1; tail -4 $O/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
( time python bench.py ) > $O/bench_n1.log 2>&1; tail -1 $O/bench_n1.log | cut -c1-400
( time python bench.py --impl reference --steps 10 --warmup 3 ) > $O/bench_ref.log 2>&1; tail -4 $O/bench_ref.log | cut -c1-300
python bench.py --steps 20 --warmup 3 --no-cpu > $O/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_c4_512.csv python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_launch.log 2>&1
tail -2 $O/launches_c4_512.csv | cut -c1-200
python bench.py --steps 12 --warmup 100 --no-cpu > $O/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_(sc|pred)' -s 400 -c 4 -o $O/prof_r1_c4_512 python bench.py --steps 12 --warmup 100 --no-cpu > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log | cut -c1-200
ls -la $O
