#!/bin/bash
# round 2l: whole-step A/B of the fused MLP kernel (bench.py --lean, 8 environments x 1024 candidates)
for o in fused_mlp=1 fused_mlp=0 fused_mlp=1 fused_mlp=0; do
  timeout 300 python bench.py --steps 40 --warmup 5 --lean --no-cpu-baseline --option $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$o', 'plans/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(d['roofline']['frac'],3), 'gemm ms', round(d['roofline']['gemm_ms_per_step'],3), 'clocks', d['clocks'])"
done
