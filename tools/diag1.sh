set -x
M3PC_FB_TRACE=1 python tools/fb_trace.py > gpurun_out/d1_fbtrace.txt 2>&1
python tools/time_breakdown.py gemm > gpurun_out/d1_gemm_microbench.txt 2>&1
for c in 1024 2048 4096; do python bench.py --workload halfcheetah_rtg_16384 --chunk $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/d1_hc16k_chunk$c.json 2>&1; done
for c in 512 2048; do python bench.py --workload walker2d_critic_1024 --chunk $c --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/d1_w1024_chunk$c.json 2>&1; done
M3PC_NO_GRAPHS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2sm --launch-skip 21 -c 21 -f -o gpurun_out/d1_gemm2sm python tools/plan_once.py walker2d_critic_1024 2 > gpurun_out/d1_ncu.log 2>&1
M3PC_NO_GRAPHS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_b1 --launch-skip 1 -c 1 -f -o gpurun_out/d1_fusedb1 python tools/plan_once.py walker2d_critic_1024 2 >> gpurun_out/d1_ncu.log 2>&1
tail -3 gpurun_out/d1_*.json
