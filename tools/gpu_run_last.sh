# scratch driver of the last measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3q; mkdir -p $O
timeout 200 python -m pytest tests/test_image_encoder.py -m gpu -q -k "pyogmaneo" > $O/t_po.log 2>&1; echo "t_po rc=$?"
tail -n 25 $O/t_po.log | cut -c1-250
