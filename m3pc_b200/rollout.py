"""The caller loops around the planning path, vectorised and pipelined.

Reference loops (one environment, one blocking plan per step):

  ReplayBuffer.online_rollout(sample_func=learner.action_sample)    finetune_omtm/replay_buffer.py:167-232
  Learner.evaluate_plan(num_episodes, episode_rtg_ref)              finetune_omtm/learner.py:648-720

Both build ``current_trajectory`` (zero-filled (1000, d) arrays + ``path_length``), call ``action_sample`` on it, clip the
action, ``env.step`` it and write action / reward back.  Here the same loop runs over E lock-step environments split into G
groups: while group g's environments step on the host, the plans of the other groups run on the GPU
(``Learner.action_sample_async`` tickets), so neither side waits for the other.  With G = 1 and E = 1 it degenerates to the
reference's loop.  Environments use the reference's (old gym) protocol: ``reset() -> obs``, ``step(a) -> (obs, r, done, info)``.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np


def new_trajectory(obs_dim: int, act_dim: int, max_path_length: int = 1000) -> Dict[str, Any]:
    """``current_trajectory`` of replay_buffer.py:192-205 / learner.py:663-675."""
    return {
        "observations": np.zeros((max_path_length, obs_dim), dtype=np.float32),
        "actions": np.zeros((max_path_length, act_dim), dtype=np.float32),
        "rewards": np.zeros((max_path_length, 1), dtype=np.float32),
        "values": np.zeros((max_path_length, 1), dtype=np.float32),
        "total_return": 0,
        "path_length": 0,
    }


def _split(n: int, groups: int) -> List[range]:
    groups = max(1, min(groups, n))
    cuts = [n * g // groups for g in range(groups + 1)]
    return [range(cuts[g], cuts[g + 1]) for g in range(groups)]


def run_episodes(learner, envs: Sequence[Any], *, rtg: Optional[Callable[[int], float]] = None, percentage: float = 1.0, plan: bool = True,
                 eval: bool = True, max_path_length: int = 1000, clip=(-1.0, 1.0), groups: int = 2,
                 on_step: Optional[Callable[[int, int, np.ndarray, float, bool], None]] = None) -> Dict[str, Any]:
    """One episode in each of ``envs`` (all reset now, stepped in lock-step).

    rtg(timestep) is the return-to-go handed to the planner at that step (``episode_rtg_ref[timestep]`` in evaluate_plan,
    learner.py:685); None -> the ``percentage`` rule of learner.py:368-385.  An environment that reports ``done`` leaves its
    group (the others keep their lock-step).  Returns per-environment ``returns`` / ``lengths`` and the trajectories.
    """
    E = len(envs)
    if E < 1:
        raise ValueError("run_episodes needs at least one environment")
    obs0 = [np.asarray(env.reset(), dtype=np.float32) for env in envs]
    obs_dim = obs0[0].shape[-1]
    act_dim = int(learner.mtm.data_shapes["actions"][1])
    trajs = [new_trajectory(obs_dim, act_dim, max_path_length) for _ in range(E)]
    for e in range(E):
        trajs[e]["observations"][0] = obs0[e]
    alive = [list(r) for r in _split(E, groups)]
    tickets: List[Any] = [None] * len(alive)
    steps = [0] * len(alive)  # timestep of each group (groups advance independently, members of a group together)

    def submit(g: int) -> None:
        if not alive[g]:
            tickets[g] = None
            return
        t = steps[g]
        r = None if rtg is None else float(rtg(t))
        tickets[g] = learner.action_sample_async([trajs[e] for e in alive[g]], percentage=percentage, plan=plan, eval=eval, rtg=r)

    for g in range(len(alive)):
        submit(g)
    while any(alive):
        for g in range(len(alive)):
            if tickets[g] is None:
                continue
            actions = np.clip(tickets[g].result(), clip[0], clip[1])  # waits for THIS group's plan only
            t = steps[g]
            still = []
            for row, e in enumerate(alive[g]):
                new_obs, reward, done, _info = envs[e].step(actions[row])
                tr = trajs[e]
                tr["actions"][t] = actions[row]
                tr["rewards"][t] = reward
                tr["total_return"] += float(reward)
                tr["path_length"] += 1
                if on_step is not None:
                    on_step(e, t, actions[row], float(reward), bool(done))
                if not done and t + 1 < max_path_length:
                    tr["observations"][t + 1] = np.asarray(new_obs, dtype=np.float32)
                    still.append(e)
            alive[g] = still
            steps[g] = t + 1
            submit(g)  # queued behind the other groups' plans; runs while the next group's environments step
    return {"returns": np.array([tr["total_return"] for tr in trajs], dtype=np.float64),
            "lengths": np.array([tr["path_length"] for tr in trajs], dtype=np.int64), "trajectories": trajs}


def evaluate_plan(learner, make_env: Callable[[], Any], num_episodes: int, episode_rtg_ref: np.ndarray, *, n_envs: Optional[int] = None,
                  groups: int = 2, max_path_length: int = 1000) -> Dict[str, float]:
    """``Learner.evaluate_plan`` (learner.py:648-720) over ``num_episodes`` episodes, ``n_envs`` at a time (default: the
    Learner's ``max_envs``): plan=True, eval=True, rtg = episode_rtg_ref[timestep], actions clipped to [-1, 1]."""
    n_envs = int(n_envs or getattr(learner, "max_envs", 1))
    rets: List[float] = []
    lens: List[int] = []
    while len(rets) < num_episodes:
        k = min(n_envs, num_episodes - len(rets))
        out = run_episodes(learner, [make_env() for _ in range(k)], rtg=lambda t: episode_rtg_ref[min(t, len(episode_rtg_ref) - 1)], plan=True,
                           eval=True, max_path_length=max_path_length, groups=groups)
        rets += out["returns"].tolist()
        lens += out["lengths"].tolist()
    return {"return_mean": float(np.mean(rets)), "return_std": float(np.std(rets)), "length_mean": float(np.mean(lens)), "episodes": len(rets)}


class LinearEnv:
    """A deterministic stand-in for the gym environments (none are installed): contracting linear dynamics with a bounded reward.
    Only for tests and the pipeline benchmark -- the reference's MuJoCo tasks are outside this path."""

    def __init__(self, obs_dim: int, act_dim: int, seed: int = 0, horizon: int = 1000, step_cost_s: float = 0.0):
        rs = np.random.RandomState(seed)
        a = rs.randn(obs_dim, obs_dim).astype(np.float32)
        self.A = (0.9 * a / max(1e-6, float(np.linalg.norm(a, 2)))).astype(np.float32)
        self.B = (0.3 * rs.randn(obs_dim, act_dim)).astype(np.float32)
        self.w = rs.randn(obs_dim).astype(np.float32) / np.sqrt(obs_dim)
        self.x0 = rs.randn(obs_dim).astype(np.float32)
        self.horizon, self.t, self.x = horizon, 0, self.x0.copy()
        self.step_cost_s = step_cost_s

    def reset(self):
        self.t, self.x = 0, self.x0.copy()
        return self.x.copy()

    def step(self, action):
        if self.step_cost_s > 0:  # emulate a physics step
            import time
            t_end = time.perf_counter() + self.step_cost_s
            while time.perf_counter() < t_end:
                pass
        self.x = (self.A @ self.x + self.B @ np.asarray(action, dtype=np.float32)).astype(np.float32)
        self.t += 1
        reward = float(np.tanh(self.w @ self.x))
        return self.x.copy(), reward, self.t >= self.horizon, {}
