#!/bin/bash
# round 2a: GPU test suite + hot-shape GEMM table (vs cuBLAS) + operand-feed / epilogue timing experiments (tuning build)
mkdir -p gpurun_out
rm -f gpurun_out/bench_size_parity.jsonl
python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^PARITY" | tail -15 > gpurun_out/r2a_pytest_gpu.txt
cat gpurun_out/r2a_pytest_gpu.txt
python tools/gemm_shapes.py --json gpurun_out/r2a_gemm_shapes.json > gpurun_out/r2a_gemm_shapes.txt 2>&1
tail -14 gpurun_out/r2a_gemm_shapes.txt
for t in 1 2 3; do
  M3PC_LIB=tuning M3PC_TUNE_GEMM=$t python tools/gemm_shapes.py --rows 106496 --no-cublas > gpurun_out/r2a_tune_gemm_$t.txt 2>&1
done
for t in 1 2 4 6 7 8; do
  M3PC_LIB=tuning M3PC_TUNE_LN=$t python tools/gemm_shapes.py --rows 106496 --no-cublas > gpurun_out/r2a_tune_ln_$t.txt 2>&1
done
grep -H "lin1\|qkv\|LN" gpurun_out/r2a_tune_*.txt | cut -c1-200
