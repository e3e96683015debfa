# chunk (rows run through the whole network before the next chunk starts) sweep at 8 lock-step environments
mkdir -p gpurun_out
for C in 1024 2048 3072 4096 8192; do
  timeout 300 python bench.py --envs 8 --chunk $C --steps 100 --no-cpu-baseline 2> gpurun_out/cs_$C.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunk', $C, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks'])" | tee -a gpurun_out/chunk_sweep.txt
done
