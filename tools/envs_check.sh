# GPU pass for the env-batched plan: parity tests, then bench at 1 / 4 / 8 / 16 lock-step environments per step.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/ec_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/ec_pytest.txt
tail -5 gpurun_out/ec_pytest.txt
for E in 1 4 8 16; do
  timeout 400 python bench.py --envs $E $( [ $E -ne 1 ] && echo --no-cpu-baseline ) > gpurun_out/ec_bench_e$E.json 2> gpurun_out/ec_bench_e$E.err
  tail -2 gpurun_out/ec_bench_e$E.err; cat gpurun_out/ec_bench_e$E.json
done
