# last gate of the round: parity tests, smoke, both bench arms (no profiler)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fc_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/fc_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/fc_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/fc_bench.json 2> gpurun_out/fc_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fc_bench_ref.json 2>&1
tail -n 3 gpurun_out/fc_pytest.txt; tail -n 2 gpurun_out/fc_smoke.txt; cut -c1-200 gpurun_out/fc_bench.json
