mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/aba_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/aba_pytest.txt
tail -8 gpurun_out/aba_pytest.txt
for S in 0 1 0 1; do
  M3PC_ATTN_SHARED=$S timeout 300 python bench.py --steps 100 --no-cpu-baseline 2> gpurun_out/aba_$S.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('attn_shared', $S, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'single p50', round(d['single_env']['p50_ms_device'],4), d['clocks'])" | tee -a gpurun_out/ab_attn.txt
  tail -2 gpurun_out/aba_$S.err
done
