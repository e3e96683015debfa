# parity tests + A/B of a GEMM scheduling switch at the default bench configuration
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ab_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ab_pytest.txt
tail -15 gpurun_out/ab_pytest.txt
for S in 0 1 0 1; do
  M3PC_GEMM_STRIDED=$S timeout 300 python bench.py --steps 100 --no-cpu-baseline 2> gpurun_out/ab_$S.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('strided', $S, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks'])" | tee -a gpurun_out/ab_strided.txt
done
