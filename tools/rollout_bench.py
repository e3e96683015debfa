#!/usr/bin/env python
"""SURVEY.md section 8f rows measured end to end: the pipelined caller loop (rollout.run_episodes) and the device-resident episode
histories (rollout.DeviceEpisodes), walker2d shapes, critic_lambda_guiding, 1024 candidates, host-side LinearEnv stand-ins.

    python tools/rollout_bench.py [steps]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_rollout import _gpu_learner  # the same Learner construction the rollout tests use
from m3pc_b200 import rollout as ro

H = int(sys.argv[1]) if len(sys.argv) > 1 else 150
for E, groups in ((8, 1), (16, 1), (16, 2), (24, 3)):
    shape, L = _gpu_learner("bf16", 1024, E)
    envs = [ro.LinearEnv(shape.obs_dim, shape.act_dim, seed=e, horizon=H) for e in range(E)]
    ro.run_episodes(L, envs, rtg=lambda t: 3.0, plan=True, eval=True, max_path_length=20, groups=groups)  # warm-up (graphs captured)
    envs = [ro.LinearEnv(shape.obs_dim, shape.act_dim, seed=e, horizon=H) for e in range(E)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = ro.run_episodes(L, envs, rtg=lambda t: 3.0, plan=True, eval=True, max_path_length=H, groups=groups)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run_episodes E={E} envs in {groups} group(s) of {E // groups}: {E * H} plans + env steps in {dt:.3f} s = {E * H / dt:.0f} plans/s end to end "
          f"(mean return {out['returns'].mean():.3f})", flush=True)
    del L
# device-resident histories: per-step upload of E*(obs+act+1) floats
E = 8
shape, L = _gpu_learner("bf16", 1024, E)
envs = [ro.LinearEnv(shape.obs_dim, shape.act_dim, seed=e, horizon=H) for e in range(E)]
for rep in range(2):
    ep = ro.DeviceEpisodes(L, n_env=E)
    obs = np.stack([env.reset() for env in envs]); ep.start(obs)
    n = 20 if rep == 0 else H
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(n):
        a = np.clip(ep.plan_async(plan=True, eval=True, rtg=3.0).result(), -1, 1)
        nxt, rew = np.zeros_like(obs), np.zeros(E, np.float32)
        for e in range(E):
            o, r, _, _ = envs[e].step(a[e]); nxt[e], rew[e] = o, r
        if t + 1 < n:
            ep.step(a, rew, nxt)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"DeviceEpisodes E={E}: {E * H} plans + env steps in {dt:.3f} s = {E * H / dt:.0f} plans/s end to end (H2D per step: {E * (shape.obs_dim + shape.act_dim + 2) * 4} B)")
