#!/bin/bash
# round 2m: attention kernels after the address-generation rewrite: parity + launch list of one default step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_sizes.py -m gpu -q -x 2>&1 | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2m_launches_default.csv python tools/plan_once.py walker2d_critic_1024 3 8 > gpurun_out/r2m.log 2>&1
python tools/launch_summary.py gpurun_out/r2m_launches_default.csv 60 | head -12
