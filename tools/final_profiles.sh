# Round-end evidence: parity tests, smoke, both bench arms, launch list, ncu --set full of every kernel of one default step.
# Only text summaries travel back (gpurun merges at most 64 MiB): the .ncu-rep is summarised on the box and deleted.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fp_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/fp_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/fp_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fp_bench_ref.json 2>&1
M3PC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/fp_launches.csv python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/fp_launches.log 2>&1
M3PC_NO_GRAPHS=1 timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/fp_step python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/fp_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/fp_step.ncu-rep > gpurun_out/fp_ncu_full_summary.txt 2>&1
rm -f /tmp/fp_step.ncu-rep
for f in gpurun_out/fp_pytest.txt gpurun_out/fp_smoke.txt gpurun_out/fp_ncu_full.log; do tail -n 3 $f; done; cut -c1-300 gpurun_out/fp_bench.json; du -sh gpurun_out
