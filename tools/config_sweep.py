#!/usr/bin/env python
"""BASELINE.json configs 4 and 5 on one GPU (device-resident, CUDA events, graph replay path):

    python tools/config_sweep.py scaled     # config 5: scaled MTM (D=1024, 8 heads, 4+2 layers, T=16, h=8), 256 .. 65536 candidates
    python tools/config_sweep.py zeroshot   # config 4: backward piid / id planners on E lock-step hopper environments
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from m3pc_b200 import synthetic as syn  # noqa: E402
from m3pc_b200.engine import engine_from_synthetic  # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def sec_scaled():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
    for model in ("scaled", "shipped"):
        w = dict(bench.WORKLOADS["scaled_rtg_4096"], model=model)
        shape = bench.model_shape(w)
        T = shape.traj_length
        h = T // 2
        g = torch.Generator(device="cuda").manual_seed(0)
        ws, wa = torch.randn(T, shape.obs_dim, device="cuda", generator=g), torch.rand(T, shape.act_dim, device="cuda", generator=g) * 2 - 1
        wr, wt = torch.randn(T, device="cuda", generator=g), torch.full((T,), 0.7, device="cuda")
        for N in (256, 1024, 4096, 16384, 65536):
            eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=N)
            seed = [0]

            def fn():
                seed[0] += 1
                eng.plan(guidance="rtg_guiding", horizon=h, n_cand=N, win_states=ws, win_actions=wa, win_rewards=wr, win_returns_tok=wt,
                         discount=0.99, temperature=0.01, lmbda=0.6, seed=seed[0])

            ms = timed(fn, reps=20 if N <= 4096 else 6)
            fl_plan, fl_row = bench.flops_per_plan(shape, N, "rtg_guiding", h)
            print(f"config5 model={model} D={shape.n_embd} T={T} h={h} N={N}: {ms:.3f} ms/plan = {1e3 / ms:.1f} plans/s = {N * 1e3 / ms / 1e6:.3f} M candidate-rollouts/s; "
                  f"dense {fl_plan / 1e12:.2f} TFLOP/plan -> {fl_plan / (ms * 1e-3) / 1e12:.0f} dense-equivalent TFLOP/s ({fl_plan / (ms * 1e-3) / 1e12 / peak:.2f} of {peak:.0f})",
                  flush=True)
            del eng
            torch.cuda.empty_cache()


def sec_zeroshot():
    shape = syn.shipped_shape("hopper")
    T, h = shape.traj_length, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    for E in (1, 32, 256, 2048):
        eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=max(E, 2))
        ws, wa = torch.randn(E, T, shape.obs_dim, device="cuda", generator=g), torch.rand(E, T, shape.act_dim, device="cuda", generator=g) * 2 - 1
        wr, wt = torch.zeros(E, T, device="cuda"), torch.full((E, T), 0.7, device="cuda")
        for mode in ("id", "piid"):
            ms = timed(lambda: eng.backward_plan(mode=mode, horizon=h, win_states=ws, win_actions=wa, win_rewards=wr, win_returns_tok=wt), reps=20)
            print(f"config4 backward {mode} E={E} envs: {ms * 1e3:.1f} us per step = {E * 1e3 / ms:.0f} env-plans/s", flush=True)
        del eng


if __name__ == "__main__":
    globals()["sec_" + sys.argv[1]]()
