#!/usr/bin/env python
"""BASELINE.json configs[3] and configs[4] at 1 / 2 / 4 / 8 GPUs of one node (one rank per GPU under torchrun; plain python = 1 GPU):

  config 4  zero-shot backward planner (piid: pi mask -> fill inferred states -> fid mask), 256 lock-step hopper environments x 512
            action draws per step, ENV-sharded (256 / world environments per GPU, no collective on the data path)
  config 5  scaled MTM (D = 1024, 8 heads, 4 + 2 layers, T = 16, h = 8), rtg_guiding, candidate count 256 .. 65 536, CANDIDATE-sharded
            (rank g owns [lo, hi); pass 1 replicated; shard records exchanged + merged inside the select kernel over NVLink peer memory)

Device-resident windows, CUDA events per step on the launching stream, max over ranks; rank 0 writes one JSON file.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config_sweep_mg.py out.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from m3pc_b200 import dist as mdist  # noqa: E402
from m3pc_b200 import synthetic as syn  # noqa: E402
from m3pc_b200.engine import engine_from_synthetic  # noqa: E402


def timed(fn, reps, warm, reduce_max, barrier):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); barrier()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        barrier()
    ts.sort()
    return reduce_max([ts[len(ts) // 2]])[0]  # median per rank, max over ranks


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    rank, local_rank, world = mdist.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)

    def reduce_max(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    res = {"world": world, "gpu": torch.cuda.get_device_name(dev), "config4": [], "config5": []}
    g = torch.Generator(device="cuda").manual_seed(rank)

    # ---- config 4 ----
    shape = syn.shipped_shape("hopper")
    T, h, E_total, C = shape.traj_length, 4, 256, 512
    E = E_total // world
    eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=max(E, 2), device=dev)
    ws, wa = torch.randn(E, T, shape.obs_dim, device=dev, generator=g), torch.rand(E, T, shape.act_dim, device=dev, generator=g) * 2 - 1
    wr, wt = torch.zeros(E, T, device=dev), torch.full((E, T), 0.7, device=dev)
    for mode in ("id", "piid"):
        seed = [0]

        def fn4():
            seed[0] += 1
            eng.backward_plan(mode=mode, horizon=h, win_states=ws, win_actions=wa, win_rewards=wr, win_returns_tok=wt, n_draws=C, seed=seed[0])

        ms = timed(fn4, 30, 5, reduce_max, barrier)
        res["config4"].append({"mode": mode, "envs_total": E_total, "envs_per_gpu": E, "draws_per_env": C, "ms_per_step": ms,
                               "env_plans_per_s": E_total * 1e3 / ms, "action_draws_per_s": E_total * C * 1e3 / ms, "collective": "none (env-sharded)"})
        if rank == 0:
            print(f"config4 {mode} world={world}: {E} envs/GPU x {C} draws: {ms * 1e3:.1f} us/step = {E_total * 1e3 / ms:.0f} env-plans/s", flush=True)
    del eng
    torch.cuda.empty_cache()

    # ---- config 5 ----
    w = dict(bench.WORKLOADS["scaled_rtg_4096"])
    shape = bench.model_shape(w)
    T = shape.traj_length
    h = T // 2
    Ns = [256, 1024, 4096, 16384, 65536]
    n_max = mdist.shard_range(max(Ns), 0, world)[1]
    eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=n_max, device=dev)
    if world > 1:
        mdist.connect_exchange(eng)
    g0 = torch.Generator(device="cuda").manual_seed(0)  # the SAME window on every rank: one plan, sharded
    ws, wa = torch.randn(T, shape.obs_dim, device=dev, generator=g0), torch.rand(T, shape.act_dim, device=dev, generator=g0) * 2 - 1
    wr, wt = torch.randn(T, device=dev, generator=g0), torch.full((T,), 0.7, device=dev)
    peak = 1376.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    for N in Ns:
        lo, hi = mdist.shard_range(N, rank, world)
        seed = [0]

        def fn5():
            seed[0] += 1
            eng.plan(guidance="rtg_guiding", horizon=h, n_cand=hi - lo, cand_offset=lo, win_states=ws, win_actions=wa, win_rewards=wr,
                     win_returns_tok=wt, discount=0.99, temperature=0.01, lmbda=0.6, seed=seed[0], exchange=world > 1)

        ms = timed(fn5, 12 if N <= 4096 else 5, 3, reduce_max, barrier)
        fl_plan, _ = bench.flops_per_plan(shape, N, "rtg_guiding", h)
        rec = {"candidates": N, "candidates_per_gpu": hi - lo, "ms_per_plan": ms, "plans_per_s": 1e3 / ms, "candidate_rollouts_per_s": N * 1e3 / ms,
               "dense_equivalent_tflops_per_gpu": fl_plan / (ms * 1e-3) / 1e12 / world, "dense_equivalent_frac_of_sustained_peak": fl_plan / (ms * 1e-3) / 1e12 / world / peak,
               "collective": "none" if world == 1 else "in-kernel peer exchange of the shard records (NVLink, CUDA IPC)"}
        res["config5"].append(rec)
        if rank == 0:
            print(f"config5 world={world} N={N}: {ms:.3f} ms/plan = {1e3 / ms:.1f} plans/s ({rec['dense_equivalent_frac_of_sustained_peak']:.2f} of peak per GPU, dense-equivalent)", flush=True)
    if world > 1:
        res["exchange_status"] = eng.exchange_status()
    if rank == 0 and out_path:
        json.dump(res, open(out_path, "w"), indent=1)
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
