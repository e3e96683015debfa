#!/bin/bash
# round 2c (gpurun --gpus 2): peer-exchange test on one GPU, then the driver-style 2-rank bench (env-parallel headline + cand_shard with
# the in-kernel NVLink exchange) and the explicit --mode cand line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "peer_exchange" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_mg2_bench.json 2> gpurun_out/r2c_mg2_bench.err
echo "rc=$?"; tail -n 5 gpurun_out/r2c_mg2_bench.err; cat gpurun_out/r2c_mg2_bench.json
timeout 300 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --mode cand --steps 30 --warmup 5 > gpurun_out/r2c_mg2_cand.json 2> gpurun_out/r2c_mg2_cand.err
echo "rc=$?"; tail -n 3 gpurun_out/r2c_mg2_cand.err; cat gpurun_out/r2c_mg2_cand.json
