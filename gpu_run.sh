cd /root/repo
mkdir -p gpurun_out
python bench.py --size 184 --steps 100 --warmup 100 --no-cpu > gpurun_out/b184.log 2>&1; python -c "
import json;d=json.loads(open('gpurun_out/b184.log').read().strip().split('\n')[-1]);print('with events ms/step',d['ms_per_step'], {k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()})"
OGN_BENCH_NOPROF=1 python bench.py --size 184 --steps 100 --warmup 100 --no-cpu 2>&1 | tail -3 | cut -c1-400
python bench.py --size 184 --steps 30 --warmup 30 --no-cpu > gpurun_out/plain184.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 200 -c 40 --csv --log-file gpurun_out/launches_184.csv python bench.py --size 184 --steps 30 --warmup 30 --no-cpu > gpurun_out/ncu184.log 2>&1
tail -2 gpurun_out/launches_184.csv | cut -c1-300
