# ncu launch list of one 8-environment step (graphs off so every kernel is listed)
set -x
mkdir -p gpurun_out
M3PC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/e8_launches.csv python bench.py --envs 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/e8_launches.log 2>&1
tail -2 gpurun_out/e8_launches.log | cut -c1-300
wc -l gpurun_out/e8_launches.csv
