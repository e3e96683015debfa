mkdir -p gpurun_out
for S in 0 1; do
  M3PC_MEGA=$S timeout 300 python bench.py --steps 100 --no-cpu-baseline 2> gpurun_out/abm_$S.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('mega', $S, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), d['clocks'])" | tee -a gpurun_out/ab_mega.txt
  tail -2 gpurun_out/abm_$S.err
done
