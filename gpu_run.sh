cd /root/repo
mkdir -p gpurun_out/final7
OGN_RUN_RANDOM_SHAPES=1 timeout 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k random_shapes_inner > gpurun_out/final7/random.log 2>&1; tail -5 gpurun_out/final7/random.log
