cd /root/repo
O=gpurun_out/r2z; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=900 run t_stage python -m pytest -m gpu -q "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[kernels=stage]" "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[kernels=stage-stage_chunks=1]"
T=600 run t_stage128 env OGN_LARGE_VARIANT=stage python -m pytest -m gpu -q tests/test_gpu_parity.py -k "oracle_size_128 or c2_video_full_size"
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mb tools/microbench_patterns.cu > $O/nvcc.log 2>&1
timeout 300 /tmp/mb > $O/microbench.jsonl 2>&1; echo "mb rc=$?"
tail -n 3 $O/t_stage.log $O/t_stage128.log
