#!/bin/bash
# Round-end evidence on ONE B200 (prefix = $1, default "r2p"): the GPU test suite, smoke, both bench arms, the ncu launch list of
# the bench command and of one eager step, and `ncu --set full` of every kernel of one default step.  Only text summaries travel
# back (gpurun merges at most 64 MiB): the .ncu-rep is summarised on the box (tools/ncu_summary.py -> tools/ncu_table.py) and deleted.
P=${1:-r2p}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/${P}_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/${P}_smoke.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err
timeout 900 python bench.py > gpurun_out/${P}_bench_default.json 2> gpurun_out/${P}_bench_default.err
# launch list of the bench command itself (graph nodes are profiled one by one) and of one eager step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${P}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --lean --no-cpu-baseline > gpurun_out/${P}_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${P}_launches_step.csv \
  python tools/plan_once.py walker2d_critic_1024 3 8 > gpurun_out/${P}_launches_step.log 2>&1
python tools/launch_summary.py gpurun_out/${P}_launches_step.csv 60 > gpurun_out/${P}_launches_step.txt 2>&1
# every kernel of one step with the full counter set
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${P}_step python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/${P}_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/${P}_step.ncu-rep gpurun_out/${P}_gemm_traffic.json > gpurun_out/${P}_step_ncu_full_raw_summary.txt 2>&1
python tools/ncu_table.py gpurun_out/${P}_step_ncu_full_raw_summary.txt > gpurun_out/${P}_step_ncu_full_summary.txt 2>&1
rm -f /tmp/${P}_step.ncu-rep
tail -n 3 gpurun_out/${P}_pytest.txt gpurun_out/${P}_smoke.txt gpurun_out/${P}_ncu_full.log; cut -c1-300 gpurun_out/${P}_bench_default.json; tail -c 400 gpurun_out/${P}_bench_reference.json
tail -n 2 gpurun_out/${P}_step_ncu_full_summary.txt; du -sh gpurun_out
