cd /root/repo
mkdir -p gpurun_out/dropin
timeout 900 python -m pytest tests/test_dropin_cpp.py -m gpu -x -q > gpurun_out/dropin/tests.log 2>&1; tail -30 gpurun_out/dropin/tests.log
