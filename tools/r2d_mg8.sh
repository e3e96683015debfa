#!/bin/bash
# round 2d (gpurun --gpus 8): the driver's own 8-rank command (env-parallel headline + cand_shard riding along)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2d_mg8_bench.json 2> gpurun_out/r2d_mg8_bench.err
echo "rc=$?"; tail -n 3 gpurun_out/r2d_mg8_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2d_mg8_bench.json").readline())
print("env-parallel", d["value"], d["e2e"]["value"], d["ms_per_step"])
print("cand_shard", json.dumps(d["cand_shard"], indent=0))
PY
