#!/bin/bash
# round 2g: fused residual GEMM + LayerNorm kernel variants (tuning build, M3PC_LN_VARIANT): 0 = round-1 kernel; template instances
# 1 = 128-row units <2 stages,4 boxes | 3,2>; 2 = 128-row <3,2>; 3 = 256-row <3,2>; 4 = 256-row <2,3>
mkdir -p gpurun_out
for v in 0 1 2 3 4; do
  echo "== variant $v"
  M3PC_LIB=tuning M3PC_LN_VARIANT=$v timeout 300 python tools/gemm_shapes.py --rows 26624 106496 --no-cublas 2>&1 | grep "LN" | tee gpurun_out/r2g_ln_variant_$v.txt
done
M3PC_LIB=tuning M3PC_LN_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_residual_layernorm_fused" 2>&1 | tail -3
M3PC_LIB=tuning M3PC_LN_VARIANT=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_residual_layernorm_fused" 2>&1 | tail -3
