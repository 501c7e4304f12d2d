cd /root/repo
O=gpurun_out/r3f; mkdir -p $O
run() { name=$1; shift; timeout "${T:-300}" "$@" > $O/$name.log 2>&1; echo "$name rc=$?" | tee -a $O/summary.txt; }
T=600 run t_ienc python -m pytest tests/test_image_encoder.py -m gpu -q
T=600 run t_stage python -m pytest -m gpu -q "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[kernels=stage]" "tests/test_gpu_parity.py::test_every_kernel_variant_and_scheduling_mode_is_bit_exact[kernels=tma]"
T=300 run ienc python tools/bench_image_encoder.py
tail -n 3 $O/t_ienc.log $O/t_stage.log; tail -n 1 $O/ienc.log | cut -c1-200
