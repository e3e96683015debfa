mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fc2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/fc2_pytest.txt
tail -n 4 gpurun_out/fc2_pytest.txt
for i in 1 2; do timeout 300 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single', round(d['single_env']['value'],1), 'p50', round(d['single_env']['p50_ms_device'],4), d['clocks']['sm_mhz'])"; done
timeout 200 python tools/time_breakdown.py envs 2>&1 | grep "E=8 \|E=1 " 
