# Multi-GPU pass (run under `gpurun --gpus N`): env-parallel and candidate-sharded bench lines at 1..N ranks, both arms at N.
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg_gpus.txt 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg_env_n1.json 2> gpurun_out/mg_env_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/mg_env_n$N.json 2> gpurun_out/mg_env_n$N.err
timeout 300 python bench.py --workload halfcheetah_rtg_16384 --mode cand --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/mg_cand_n1.json 2> gpurun_out/mg_cand_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload halfcheetah_rtg_16384 --mode cand --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/mg_cand_n$N.json 2> gpurun_out/mg_cand_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/mg_ref_n$N.json 2> gpurun_out/mg_ref_n$N.err
tail -n 2 gpurun_out/mg_*.json; tail -n 5 gpurun_out/mg_*.err
