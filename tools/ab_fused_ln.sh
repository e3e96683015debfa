mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/abl_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/abl_pytest.txt
tail -12 gpurun_out/abl_pytest.txt
run() { timeout 300 python bench.py --steps 100 --no-cpu-baseline "$@" 2> gpurun_out/abl.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$TAG', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'p50', round(d['single_env']['p50_ms_device'],4), 'gemm_ms', round(d['roofline']['gemm_ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_fused_ln.txt; }
TAG=unfused M3PC_NO_FUSED_LN=1 run
TAG=fused run
TAG=unfused M3PC_NO_FUSED_LN=1 run
TAG=fused run
tail -3 gpurun_out/abl.err
