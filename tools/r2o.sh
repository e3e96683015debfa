#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gemm_fp32" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -5
python tools/sgemm_time.py 2>&1 | tee gpurun_out/r2o_sgemm.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "fp32" 2>&1 | tail -3
timeout 900 python bench.py --precision fp32 --steps 6 --warmup 3 --lean --no-cpu-baseline > gpurun_out/r2o_bench_fp32.json 2> gpurun_out/r2o_bench_fp32.err; tail -2 gpurun_out/r2o_bench_fp32.err; cut -c1-200 gpurun_out/r2o_bench_fp32.json
