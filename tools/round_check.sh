# One GPU pass over the round's gates: parity tests, smoke, bench (both arms), launch list of the default step.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/rc_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/rc_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/rc_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/rc_bench.json 2> gpurun_out/rc_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/rc_bench_ref.json 2>&1
M3PC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/rc_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/rc_launches.log 2>&1
tail -3 gpurun_out/rc_pytest.txt gpurun_out/rc_smoke.txt; cat gpurun_out/rc_bench.json; tail -c 600 gpurun_out/rc_bench_ref.json
