#!/bin/bash
# round 2e: gemm_ln128 (128-row units, double-buffered accumulators): kernel-level parity, then A/B against the 256-row kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_residual_layernorm_fused" 2>&1 | tail -15
for rows in 128 256; do
  M3PC_LIB=tuning M3PC_LN_UNIT_ROWS=$rows timeout 300 python tools/gemm_shapes.py --rows 13312 26624 106496 --no-cublas 2>&1 | grep "LN" > gpurun_out/r2e_ln_unit_$rows.txt
  echo "== unit rows $rows"; cat gpurun_out/r2e_ln_unit_$rows.txt
done
for t in 1 6 7 8; do
  echo "== unit rows 128, M3PC_TUNE_LN=$t"; M3PC_LIB=tuning M3PC_TUNE_LN=$t timeout 300 python tools/gemm_shapes.py --rows 106496 --no-cublas 2>&1 | grep "LN"
done
