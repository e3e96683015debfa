#!/usr/bin/env python
"""Run a few device-resident plans of a bench workload, eagerly (option "graphs" = 0, so ncu sees every kernel), for ncu captures:
    python tools/plan_once.py [workload] [n_plans] [n_env] [n_cand override]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from m3pc_b200 import synthetic as syn
from m3pc_b200.engine import engine_from_synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "walker2d_critic_1024"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
E = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = dict(bench.WORKLOADS[name])
if len(sys.argv) > 4:
    w["n_cand"] = int(sys.argv[4])
shape = bench.model_shape(w)
crit = w["guidance"] != "rtg_guiding"
eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=w["n_cand"] * E,
                            critic_sd=syn.make_critic_state_dict(shape) if crit else None, obs_norm=syn.make_obs_norm(shape) if crit else None)
eng.set_option("graphs", 0)
T = shape.traj_length
g = torch.Generator(device="cuda").manual_seed(0)
lead = (E,) if E > 1 else ()
ws, wa = torch.randn(*lead, T, shape.obs_dim, device="cuda", generator=g), torch.rand(*lead, T, shape.act_dim, device="cuda", generator=g) * 2 - 1
wr, wt = torch.randn(*lead, T, device="cuda", generator=g), torch.full((*lead, T), 0.7, device="cuda")
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # ncu --profile-from-start off: only the last plan is captured
    eng.plan(guidance=w["guidance"], horizon=4, n_cand=w["n_cand"], win_states=ws, win_actions=wa, win_rewards=wr, win_returns_tok=wt,
             discount=0.99, temperature=w["temperature"], lmbda=0.6, seed=i, n_env=E)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches per plan", eng.last_launch_count(), "last ms", eng.last_device_ms())
