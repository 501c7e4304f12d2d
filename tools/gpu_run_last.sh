# scratch driver of the last measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3p; mkdir -p $O
timeout 300 env OGN_RUN_RANDOM_SHAPES=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -k random_shapes_inner > $O/t_inner.log 2>&1; echo "t_inner rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > $O/t_all.log 2>&1; echo "t_all rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
tail -n 2 $O/t_inner.log; tail -n 3 $O/t_all.log; tail -n 1 $O/smoke.log
