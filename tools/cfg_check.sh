# parity tests, then BASELINE configs 4 (zero-shot, E envs) and 5 (scaled model sweep) on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/cc_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/cc_pytest.txt
tail -12 gpurun_out/cc_pytest.txt
timeout 600 python tools/config_sweep.py zeroshot > gpurun_out/cc_zeroshot.txt 2>&1; cat gpurun_out/cc_zeroshot.txt | tail -12
timeout 900 python tools/config_sweep.py scaled > gpurun_out/cc_scaled.txt 2>&1; cat gpurun_out/cc_scaled.txt | tail -14
