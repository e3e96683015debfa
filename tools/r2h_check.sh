#!/bin/bash
# round 2h: full GPU suite with the template LN kernel + decoder fusions, then the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2h_pytest_gpu.txt
timeout 600 python bench.py --steps 30 --warmup 5 --lean > gpurun_out/r2h_bench_lean.json 2> gpurun_out/r2h_bench.err; tail -2 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h_bench_lean.json").readline())
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "single", d["single_env"]["value"], d["single_env"]["p50_ms_device"], "roofline", d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], "launches", d["gpu_launches"]/d["steps"])
PY
