# scratch driver of the last measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3n; mkdir -p $O
timeout 600 python -m pytest tests/test_image_encoder.py -m gpu -q > $O/t_ienc.log 2>&1; echo "t_ienc rc=$?"
timeout 600 env OGN_IENC_TILE=0 python -m pytest tests/test_image_encoder.py -m gpu -q > $O/t_ienc_twopass.log 2>&1; echo "t_ienc_twopass rc=$?"
tail -n 3 $O/t_ienc.log $O/t_ienc_twopass.log
