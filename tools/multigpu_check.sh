#!/bin/bash
# Multi-GPU pass (run under `gpurun --gpus N`, N = 2 / 4 / 8): the driver's own N-rank command.  One bench line carries the
# env-parallel headline (config 2, no collective) and, under "cand_shard", the candidate-sharded config 3 (halfcheetah rtg 16 384)
# with the in-kernel NVLink record exchange and its 1-rank point measured in the same run.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err
echo "rc=$?"; tail -n 3 gpurun_out/mg${N}_bench.err
timeout 300 $TR --nproc-per-node $N --master-port 29542 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/mg${N}_ref.json 2> gpurun_out/mg${N}_ref.err
python - "$N" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/mg{sys.argv[1]}_bench.json").readline())
print("env-parallel", d["value"], d["e2e"]["value"], d["ms_per_step"])
print("cand_shard", json.dumps(d.get("cand_shard"), indent=0))
PY
