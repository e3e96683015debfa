# 8-GPU pass (run under `gpurun --gpus 8`): candidate-sharded config 3 at 4 and 8 ranks, env-parallel default bench at 8 ranks.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg8_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/mg8_env_n8.json 2> gpurun_out/mg8_env_n8.err
timeout 240 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --workload halfcheetah_rtg_16384 --mode cand --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/mg8_cand_n8.json 2> gpurun_out/mg8_cand_n8.err
timeout 240 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --workload halfcheetah_rtg_16384 --mode cand --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/mg8_cand_n4.json 2> gpurun_out/mg8_cand_n4.err
timeout 240 python bench.py --workload halfcheetah_rtg_16384 --mode cand --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/mg8_cand_n1.json 2> gpurun_out/mg8_cand_n1.err
for f in gpurun_out/mg8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline()); print(sys.argv[1], d["n_gpus"], d["config"]["parallelism"], "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],3), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
tail -n 3 gpurun_out/mg8_*.err
