#!/bin/bash
# round 2j: ncu --set full of every kernel of ONE default step (summary only travels back) + a source-level capture of the attention kernels
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2j_step python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/r2j_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r2j_step.ncu-rep gpurun_out/r2j_traffic.json > gpurun_out/r2j_step_ncu_full_raw_summary.txt 2>&1
tail -3 gpurun_out/r2j_ncu_full.log
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:attention_mma -f -o gpurun_out/r2j_attention python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/r2j_ncu_att.log 2>&1
ls -la gpurun_out/r2j_attention.ncu-rep /tmp/r2j_step.ncu-rep
rm -f /tmp/r2j_step.ncu-rep
