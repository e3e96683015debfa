mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "residual_layernorm_fused or env_batched_plan_rows" 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 100 --no-cpu-baseline "$@" 2> gpurun_out/abl.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$TAG', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'p50', round(d['single_env']['p50_ms_device'],4), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_fused_ln2.txt; }
TAG=unfused M3PC_NO_FUSED_LN=1 run
TAG=fused_l2pf run
TAG=unfused M3PC_NO_FUSED_LN=1 run
TAG=fused_l2pf run
