# Round-end evidence: parity tests, smoke, both bench arms, launch list, ncu --set full of the GEMM and of the memory-bound kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/fp_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/fp_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/fp_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/fp_bench.json 2> gpurun_out/fp_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/fp_bench_ref.json 2>&1
M3PC_NO_GRAPHS=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/fp_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/fp_launches.log 2>&1
M3PC_NO_GRAPHS=1 timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_2sm --launch-skip 28 --launch-count 27 \
  -f -o gpurun_out/fp_gemm2sm python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/fp_gemm2sm.log 2>&1
M3PC_NO_GRAPHS=1 timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:embed|layernorm|attention|rowdot|fill_rows|critic|score|select|candidates' --launch-skip 34 --launch-count 33 \
  -f -o gpurun_out/fp_memkernels python tools/plan_once.py walker2d_critic_1024 2 8 > gpurun_out/fp_memkernels.log 2>&1
tail -3 gpurun_out/fp_pytest.txt gpurun_out/fp_smoke.txt gpurun_out/fp_gemm2sm.log gpurun_out/fp_memkernels.log; cut -c1-400 gpurun_out/fp_bench.json
