mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/em_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/em_pytest.txt
tail -n 3 gpurun_out/em_pytest.txt
M3PC_NO_GRAPHS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:embed --csv --log-file gpurun_out/em.csv python tools/plan_once.py walker2d_critic_1024 2 8 > /dev/null 2>&1
grep -v "^==" gpurun_out/em.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin): print(r['Kernel Name'][:50], r['Grid Size'], r['Metric Value'], r['Metric Unit'])"
timeout 300 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],3), 'single p50', round(d['single_env']['p50_ms_device'],4), d['clocks']['sm_mhz'])"
