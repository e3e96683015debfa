# ncu --set full on the tensor-core GEMM launches of pass 2 of one 8-environment step (second plan; graphs off)
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2sm --launch-skip 26 --launch-count 25 \
  -f -o gpurun_out/e8_gemm2sm python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/e8_gemm2sm.log 2>&1
tail -3 gpurun_out/e8_gemm2sm.log
ls -la gpurun_out/e8_gemm2sm.ncu-rep
