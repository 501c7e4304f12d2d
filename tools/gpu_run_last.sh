# scratch driver of the last single-GPU measurement call of round 2 (run from the repo root: bash tools/gpu_run_last.sh); the
# commands behind every committed artifact are in profiles/commands.md
cd /root/repo
O=gpurun_out/r3j; mkdir -p $O
CMD="python bench.py --no-cpu --steps 10 --warmup 5 --burnin 100"
$CMD > $O/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_(sc|pred)' -s 400 -c 4 -o $O/prof_c4_512_final $CMD > $O/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -n 2 $O/ncu_full.log | cut -c1-300
