#!/bin/bash
# round 2f: launch lists (ncu gpu__time_duration, serialised) of the 8-GPU candidate shard shape (2048 candidates, one window) and
# of the default bench step
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_cheetah_2048.csv python tools/plan_once.py halfcheetah_rtg_16384 3 1 2048 > gpurun_out/r2f_a.log 2>&1
tail -2 gpurun_out/r2f_a.log
python tools/launch_summary.py gpurun_out/r2f_launches_cheetah_2048.csv 31 -v > gpurun_out/r2f_launches_cheetah_2048.txt; cat gpurun_out/r2f_launches_cheetah_2048.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches_default.csv python tools/plan_once.py walker2d_critic_1024 3 8 > gpurun_out/r2f_b.log 2>&1
tail -2 gpurun_out/r2f_b.log
python tools/launch_summary.py gpurun_out/r2f_launches_default.csv 61 > gpurun_out/r2f_launches_default.txt; head -24 gpurun_out/r2f_launches_default.txt
