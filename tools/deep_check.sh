mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scaled or forward" > gpurun_out/dc_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/dc_pytest.txt
tail -15 gpurun_out/dc_pytest.txt
timeout 900 python tools/config_sweep.py scaled 2>&1 | grep "model=scaled" | tee gpurun_out/dc_scaled_restricted.txt
M3PC_DEC_FULL=1 timeout 900 python tools/config_sweep.py scaled 2>&1 | grep "model=scaled" | tee gpurun_out/dc_scaled_full.txt
