# ncu --set full on the memory-bound fused kernels (everything that is not a tensor-core GEMM) of one 8-environment step
set -x
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:embed|layernorm|attention|rowdot|fill_rows|critic|score|select|candidates' --launch-skip 33 --launch-count 32 \
  -f -o gpurun_out/e8_memkernels python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/e8_memkernels.log 2>&1
tail -3 gpurun_out/e8_memkernels.log
ls -la gpurun_out/e8_memkernels.ncu-rep
