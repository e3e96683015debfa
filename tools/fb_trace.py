import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench
from m3pc_b200 import synthetic as syn
from m3pc_b200.engine import engine_from_synthetic
w = bench.WORKLOADS["hopper_rtg_1024"]; shape = bench.model_shape(w)
eng = engine_from_synthetic(shape, syn.make_state_dict(shape, 0), syn.make_tokenizer_stats(shape, 1), precision="bf16", max_batch=4)
T = shape.traj_length
ws, wa = torch.randn(T, shape.obs_dim, device="cuda"), torch.rand(T, shape.act_dim, device="cuda") * 2 - 1
wr, wt = torch.randn(T, device="cuda"), torch.full((T,), 0.7, device="cuda")
for i in range(2):
    print("--- call", i, flush=True)
    eng.plan(guidance="mtm_sampling", horizon=4, n_cand=1, win_states=ws, win_actions=wa, win_rewards=wr, win_returns_tok=wt, discount=0.99, temperature=1.0, lmbda=0.6, seed=i)
    torch.cuda.synchronize()
