#!/bin/bash
# round 2n: attention parity after the plain-kernel rewrite, launch list, fp32-mode bench (reference-grade precision path)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2n_launches_default.csv python tools/plan_once.py walker2d_critic_1024 3 8 > gpurun_out/r2n.log 2>&1
python tools/launch_summary.py gpurun_out/r2n_launches_default.csv 60 2>/dev/null | grep -E "attention|total"
timeout 900 python bench.py --precision fp32 --steps 6 --warmup 3 --lean --no-cpu-baseline > gpurun_out/r2n_bench_fp32.json 2> gpurun_out/r2n_bench_fp32.err; tail -2 gpurun_out/r2n_bench_fp32.err; cut -c1-1500 gpurun_out/r2n_bench_fp32.json
