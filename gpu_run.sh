cd /root/repo
mkdir -p gpurun_out/final2
O=gpurun_out/final2
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/tests_gpu.log 2>&1; tail -6 $O/tests_gpu.log | head -3
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
show() { grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.1f ms/step %.4f e2e %.1f step_frac %.3f host_enqueue_ms %.4f cpu %s'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac'],d['host_enqueue_ms_per_step'],(d.get('cpu_baseline') or {}).get('value')))"; }
for w in c1 c2 c3; do echo $w; timeout 600 python bench.py --workload $w 2>&1 | tee $O/bench_$w.log | show; done
echo c5; timeout 900 python bench.py --workload c5 2>&1 | tee $O/bench_c5.log | show
echo c4; ( time timeout 900 python bench.py ) 2>&1 | tee $O/bench_c4_n1.log | show
echo ref; ( time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 ) 2>&1 | tee $O/bench_ref.log | grep -a '^{' | cut -c1-200
timeout 900 python bench.py --no-cpu > $O/plain_launch.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1202 -c 400 --csv --log-file $O/launches_c4_512_steady.csv python bench.py --no-cpu > $O/ncu_launch.log 2>&1
tail -2 $O/launches_c4_512_steady.csv | cut -c1-160
