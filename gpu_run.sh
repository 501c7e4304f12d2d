cd /root/repo
mkdir -p gpurun_out/final5
timeout 900 python -m pytest tests/test_gpu_features.py -m gpu -x -q -k "peer" > gpurun_out/final5/peer.log 2>&1; tail -15 gpurun_out/final5/peer.log
