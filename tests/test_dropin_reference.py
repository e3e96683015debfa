"""The drop-in of INTEGRATION.md section 3, EXECUTED: ``PlannerMixin`` mixed into the UNMODIFIED reference ``Learner``
(research/finetune_omtm/learner.py:17, staged under oracle/_ref by oracle/stage_ref.py), whose ``self.mtm`` stays the
reference's own trainable ``omtm`` and whose ``self.iql.qf`` stays the reference's ``TwinQ``.

  * planning calls (``action_sample`` as ``ReplayBuffer.online_rollout`` and ``Learner.evaluate_plan`` make them,
    replay_buffer.py:208-216, learner.py:683-689) run on the B200 engine and reproduce the reference's own actions
    (golden fixture written by the reference + a live CPU twin of the same Learner);
  * training keeps working on the reference module (forward with gradients + optimiser step), and the next plan uses the
    updated weights without any ``mark_dirty`` call (parameter version counters).
"""
import json
import os

import numpy as np
import pytest
import torch

from m3pc_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _ref():
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("the reference is not staged (python oracle/stage_ref.py where /root/reference exists)")
    return rh


def _mix(L):
    from m3pc_b200.learner import PlannerMixin
    L.__class__ = type("FastLearner", (PlannerMixin, type(L)), {})  # mixin FIRST: its planners shadow the reference's
    return L


def _noise(seed, n, T, A):
    torch.manual_seed(seed)
    eps = torch.randn(n, 1, T, 1, A)
    q = torch.empty(n).exponential_(1)
    return eps, q


def test_mixin_on_the_reference_learner_reproduces_golden_actions(golden_dir):
    rh = _ref()
    z = np.load(os.path.join(golden_dir, "planner.npz"))
    meta = json.loads(str(z["meta"]))
    done = 0
    for case in meta["cases"]:
        if not case["plan"] or case["guidance"] == "noise_adding_lambda":
            continue
        shape = syn.shipped_shape(case["env"])
        L = _mix(rh.build_learner(shape, guidance=case["guidance"], n_cand=case["n_cand"], temperature=case["temperature"], device="cuda"))
        assert type(L.mtm).__module__.startswith("research.")  # still the reference's trainable module
        T, h = shape.traj_length, case["horizon"]
        eps = torch.from_numpy(z[f"{case['tag']}/eps"])
        L.injected_noise = (eps[:, 0, T - h:, 0, :].contiguous().cuda(), torch.from_numpy(z[f"{case['tag']}/q"]).cuda())
        hist = syn.make_history(shape, seed=case["hist_seed"], path_length=case["path_length"])
        act = L.action_sample(hist, percentage=case["percentage"], plan=True, eval=case["eval"], rtg=case["rtg"])
        ref = z[f"{case['tag']}/action"]
        assert tuple(act.shape) == ref.shape and act.is_cuda
        if case["eval"]:
            np.testing.assert_allclose(act.cpu().numpy(), ref, rtol=0, atol=2e-2, err_msg=case["tag"])
            done += 1
    assert done >= 5


def test_training_continues_on_the_reference_module_and_plans_follow_it():
    rh = _ref()
    shape = syn.shipped_shape("walker2d")
    N, temp, guidance = 96, 1.0, "critic_lambda_guiding"
    T, A, h = shape.traj_length, shape.act_dim, 4
    G = _mix(rh.build_learner(shape, guidance=guidance, n_cand=N, temperature=temp, device="cuda"))
    Cpu = rh.build_learner(shape, guidance=guidance, n_cand=N, temperature=temp, device="cpu")  # the reference, untouched, as the truth
    hist = syn.make_history(shape, seed=4, path_length=80)
    eps, q = _noise(7, N, T, A)
    G.injected_noise = (eps[:, 0, T - h:, 0, :].contiguous().cuda(), q.cuda())

    def truth():
        torch.manual_seed(7)
        return Cpu.action_sample(hist, plan=True, eval=True, rtg=3.0)

    a0 = G.action_sample(hist, plan=True, eval=True, rtg=3.0).cpu()
    np.testing.assert_allclose(a0.numpy(), truth().numpy(), rtol=0, atol=2e-2)
    # the call shape of ReplayBuffer.online_rollout (replay_buffer.py:208-216): exploration action, then .cpu().numpy()
    s = G.action_sample(hist, percentage=1.0, plan=True)
    assert s.shape == (1, A) and np.isfinite(s.cpu().numpy()).all()
    # one optimiser step on the REFERENCE modules (gradients flow through the reference's own forward)
    from m3pc_b200 import masks as M
    traj = {k: torch.from_numpy(v).cuda() for k, v in syn.make_trajectories(shape, 4, 5).items()}
    enc = G.tokenizer_manager.encode(traj)
    G.mtm.train()
    opt = torch.optim.Adam(G.mtm.parameters(), lr=3e-3)
    out = G.mtm(enc, M.create_rcbc_mask(T, "cuda", 4))
    loss = sum(out[k].pow(2).mean() for k in ("states", "rewards", "returns")) + out["actions"].mean.pow(2).mean()
    loss.backward()
    opt.step()
    G.mtm.eval()
    qopt = torch.optim.Adam(G.iql.qf.parameters(), lr=3e-3)
    G.iql.qf(torch.randn(8, shape.obs_dim, device="cuda"), torch.rand(8, A, device="cuda")).mean().backward()
    qopt.step()
    Cpu.mtm.load_state_dict({k: v.cpu() for k, v in G.mtm.state_dict().items()})
    Cpu.iql.qf.load_state_dict({k: v.cpu() for k, v in G.iql.qf.state_dict().items()})
    a1 = G.action_sample(hist, plan=True, eval=True, rtg=3.0).cpu()  # no mark_dirty: version counters moved
    t1 = truth()
    assert float((a1 - a0).abs().max()) > 1e-3, "the optimiser steps did not reach the planner"
    np.testing.assert_allclose(a1.numpy(), t1.numpy(), rtol=0, atol=2e-2)
