mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/ab2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ab2_pytest.txt
tail -6 gpurun_out/ab2_pytest.txt
run() { python bench.py --steps 100 --no-cpu-baseline "$@" 2> gpurun_out/ab2.err | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$TAG', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'p50', round(d['single_env']['p50_ms_device'],4), d['clocks']['sm_mhz'])" | tee -a gpurun_out/ab_attn2.txt; }
TAG=shared0 M3PC_ATTN_SHARED=0 run
TAG=shared1 run
TAG=shared0 M3PC_ATTN_SHARED=0 run
TAG=shared1 run
TAG=envs16 run --envs 16
TAG=envs4 run --envs 4
