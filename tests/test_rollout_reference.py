"""The caller loop of the path, checked against the REFERENCE's own loop.

``ReplayBuffer.online_rollout`` (research/finetune_omtm/replay_buffer.py:167-232, unmodified, staged under oracle/_ref) is the
code that calls ``Learner.action_sample`` in production.  Here it drives the B200 planner on a deterministic stand-in
environment (no MuJoCo in this image), and the episode it collects must be reproduced step for step by
``m3pc_b200.rollout.run_episodes`` -- the pipelined loop with asynchronous action read-back -- given the same planner state.
"""
from collections import deque, namedtuple
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from m3pc_b200 import rollout as ro
from m3pc_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _planner(n_cand=96):
    from m3pc_b200.learner import Learner
    from m3pc_b200.mtm_model import omtmConfig
    from m3pc_b200.tokenizers import manager_from_stats
    shape = syn.shipped_shape("hopper")
    cfg = SimpleNamespace(traj_length=shape.traj_length, device="cuda", action_samples=n_cand, discount=0.99, temperature=0.01, horizon=4,
                          plan_guidance="rtg_guiding", lmbda=0.6)
    mcfg = omtmConfig(n_embd=shape.n_embd, n_head=shape.n_head, n_enc_layer=shape.n_enc_layer, n_dec_layer=shape.n_dec_layer, dropout=0.1,
                      norm="none", precision="bf16", max_batch=n_cand)
    L = Learner(cfg, None, shape.data_shapes, mcfg, None, None, None, manager_from_stats(syn.make_tokenizer_stats(shape, 1)),
                {k: False for k in shape.data_shapes})
    L.mtm.load_state_dict({k: torch.from_numpy(v) for k, v in syn.make_state_dict(shape, 0).items()})
    L.seed = 5
    return shape, L


def _reference_buffer(env, obs_dim, act_dim, max_path_length):
    from oracle import ref_harness as rh
    if not rh.available():
        pytest.skip("the reference is not staged (python oracle/stage_ref.py where /root/reference exists)")
    rh._install()
    from research.finetune_omtm.replay_buffer import ReplayBuffer
    B = object.__new__(ReplayBuffer)  # __init__ wants a D4RL dataset; online_rollout only needs the fields below
    B.env = env
    B.max_path_length, B.observation_dim, B.action_dim = max_path_length, obs_dim, act_dim
    B.traj_buffer_size = 4
    B.cfg = SimpleNamespace(rtg_percent=0.9, plan=True, clip_min=-1.0, clip_max=1.0)
    B.experience = namedtuple("Experience", field_names=["state", "action", "reward", "next_state", "done"])
    B.online_trans_buffer = deque(maxlen=10000)
    B.discounts = (0.99 ** np.arange(max_path_length))[:, None]
    B.use_avg = False
    B.p_length_list, B.p_return_list, B.total_step = [], [], 0
    collected = []
    B.update_buffer = lambda trajs: collected.extend(trajs)  # the dataset-side bookkeeping is not on this path
    return B, collected


@pytest.mark.parametrize("horizon", [12, 25])
def test_reference_online_rollout_equals_run_episodes(horizon):
    shape, L = _planner()
    obs, A = shape.obs_dim, shape.act_dim
    env = ro.LinearEnv(obs, A, seed=3, horizon=horizon)
    B, collected = _reference_buffer(env, obs, A, max_path_length=40)
    L.__dict__["_plan_counter"] = 0
    L._engine()
    L.__dict__["_plan_counter"] = 0  # same Philox keys for both loops
    log = B.online_rollout(L.action_sample)  # the reference's loop: blocking .cpu() per step, exploration action (eval=False)
    assert len(collected) == 1 and collected[0]["path_length"] == horizon
    ref = collected[0]
    assert abs(log["explore/rollout_return_mean"] - float(ref["rewards"].sum())) < 1e-5
    L.__dict__["_plan_counter"] = 0
    out = ro.run_episodes(L, [ro.LinearEnv(obs, A, seed=3, horizon=horizon)], rtg=None, percentage=0.9, plan=True, eval=False, max_path_length=40, groups=1)
    tr = out["trajectories"][0]
    assert tr["path_length"] == horizon
    np.testing.assert_array_equal(tr["actions"][:horizon], ref["actions"][:horizon])
    np.testing.assert_array_equal(tr["observations"][:horizon], ref["observations"][:horizon])
    np.testing.assert_allclose(tr["rewards"][:horizon], ref["rewards"][:horizon], rtol=0, atol=0)
    assert float(np.abs(ref["actions"][:horizon]).max()) <= 1.0 and float(np.abs(ref["actions"][:horizon]).std()) > 1e-3
