cd /root/repo
mkdir -p gpurun_out/graphs
O=gpurun_out/graphs
timeout 1200 python -m pytest tests -m gpu -x -q > $O/tests.log 2>&1; tail -5 $O/tests.log
show() { grep -a '^{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('steps/s %.1f ms/step %.4f e2e %.1f host_enqueue_ms %.4f launches %d'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['host_enqueue_ms_per_step'],d['gpu_launches']))"; }
for g in 1 0; do
for w in c1 c2 c3; do echo "graphs=$g $w"; OGN_GRAPHS=$g timeout 600 python bench.py --workload $w --no-cpu 2>&1 | show; done
echo "graphs=$g c5x64"; OGN_GRAPHS=$g timeout 600 python bench.py --workload c5 --ensemble 64 --no-cpu 2>&1 | show
echo "graphs=$g c4@184"; OGN_GRAPHS=$g timeout 600 python bench.py --size 184 --no-cpu 2>&1 | show
done
echo "graphs=1 c4"; timeout 600 python bench.py --no-cpu 2>&1 | show
