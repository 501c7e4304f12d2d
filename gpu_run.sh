cd /root/repo
mkdir -p gpurun_out/ckpt
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "checkpoint or pyogmaneo" > gpurun_out/ckpt/tests.log 2>&1; tail -25 gpurun_out/ckpt/tests.log
