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

from types import SimpleNamespace
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


class DeviceEpisodes:
    """Device-resident histories of E lock-step episodes (``m3pc_ring_*`` in include/m3pc.h).

    The reference keeps each episode as zero-filled (1000, d) numpy arrays and re-slices a window out of them on the host at
    every step (learner.py:346-366).  Here the arrays live in HBM: a step uploads only the new observation and the action /
    reward just taken (E * (obs + act + 1) floats), and the planner's (E, T, .) windows are cut on the device.

        ep = DeviceEpisodes(learner, n_env=E); ep.start(obs0)                     # obs0: (E, obs)
        a = ep.plan_async(rtg=3.0, plan=True, eval=True).result()                 # (E, act)
        ep.step(a, rewards, next_obs)                                             # then plan again
    """

    def __init__(self, learner, n_env: int, max_path_length: int = 1000):
        import ctypes as C

        import torch

        from . import _native as nat
        self.learner, self.E, self.L = learner, int(n_env), int(max_path_length)
        if self.E < 1 or self.E > int(getattr(learner, "max_envs", 1)):
            raise ValueError(f"{n_env} environments but the Learner was built with max_envs={getattr(learner, 'max_envs', 1)}")
        learner._engine()
        self._nat, self._C, self._torch = nat, C, torch
        self.obs = int(learner.mtm.data_shapes["states"][1])
        self.act = int(learner.mtm.data_shapes["actions"][1])
        self.T = int(learner.cfg.traj_length)
        dev = learner.mtm.pos_embed.device
        w = self.obs + self.act + 1
        self.ring = torch.zeros(self.E, self.L, w, device=dev)
        self.t = -1  # path_length of the current step; -1 before start()
        self._stage = [SimpleNamespace(host=torch.zeros(self.E * (w + 1)).pin_memory(), event=torch.cuda.Event(), used=False) for _ in range(2)]
        self._turn = 0
        self._dev_stage = torch.zeros(self.E * (w + 1), device=dev)
        E, T = self.E, self.T
        self.win_states = torch.zeros(E, T, self.obs, device=dev)
        self.win_actions = torch.zeros(E, T, self.act, device=dev)
        self.win_rewards = torch.zeros(E, T, device=dev)
        self.win_returns = torch.zeros(E, T, device=dev)

    def _slot(self):
        s = self._stage[self._turn]
        self._turn ^= 1
        if s.used:
            s.event.synchronize()
        return s

    def _upload(self, slot, n: int):
        self._dev_stage[:n].copy_(slot.host[:n], non_blocking=True)
        slot.event.record()
        slot.used = True

    def _stream(self):
        return self._torch.cuda.current_stream().cuda_stream

    def start(self, obs0) -> None:
        """Begin E episodes: clears the histories and stores the first observations (E, obs)."""
        self.ring.zero_()
        self.t = 0
        self._append(np.asarray(obs0, dtype=np.float32), None, None)

    def step(self, actions, rewards, next_obs) -> None:
        """Record the actions taken at the current step and their rewards, and the observations of the next step."""
        if self.t < 0:
            raise RuntimeError("DeviceEpisodes.step() before start()")
        if self.t + 1 >= self.L:
            raise RuntimeError("episode longer than max_path_length")
        self.t += 1
        self._append(np.asarray(next_obs, dtype=np.float32), np.asarray(actions, dtype=np.float32), np.asarray(rewards, dtype=np.float32))

    def _append(self, obs, act, rew) -> None:
        E, o, a = self.E, self.obs, self.act
        if obs.shape != (E, o):
            raise ValueError(f"observations have shape {obs.shape}, expected {(E, o)}")
        slot = self._slot()
        h = slot.host.numpy()
        h[:E * o] = obs.reshape(-1)
        n = E * o
        pa = pr = None
        if act is not None:
            h[n:n + E * a] = act.reshape(E * a)
            h[n + E * a:n + E * a + E] = rew.reshape(E)
            pa = self._dev_stage.data_ptr() + 4 * n
            pr = pa + 4 * E * a
            n += E * a + E
        self._upload(slot, n)
        nat = self._nat
        nat.check(nat.lib().m3pc_ring_append(self.ring.data_ptr(), E, self.L, o, a, self.t, self._dev_stage.data_ptr(), pa, pr, self._stream()),
                  "m3pc_ring_append")

    def plan_async(self, percentage: float = 1.0, plan: bool = True, eval: bool = False, rtg=None):
        """``Learner.action_sample_async`` on the device-resident histories: returns a ticket for the (E, act) actions."""
        L = self.learner
        if eval == True:  # noqa: E712
            assert rtg is not None
        if self.t < 0:
            raise RuntimeError("DeviceEpisodes.plan_async() before start()")
        E, T = self.E, self.T
        horizon = L._clamped_horizon({"path_length": self.t})
        if rtg is not None:
            r = np.broadcast_to(np.asarray(rtg, dtype=np.float64).reshape(-1), (E,)) if np.ndim(rtg) else np.full(E, float(rtg))
        else:
            stats = L.tokenizer_manager.tokenizers["returns"].stats
            r = np.full(E, float(np.asarray(stats.min + (stats.max - stats.min) * percentage).reshape(-1)[0]))
        slot = self._slot()
        slot.host.numpy()[:E] = L._returns_tok(r * np.ones(E)).reshape(E)
        self._upload(slot, E)
        nat = self._nat
        nat.check(nat.lib().m3pc_ring_windows(self.ring.data_ptr(), E, self.L, self.obs, self.act, self.t, horizon, T,
                                              1 if getattr(L, "_future_obs_windows", False) else 0, self._dev_stage.data_ptr(),
                                              self.win_states.data_ptr(), self.win_actions.data_ptr(), self.win_rewards.data_ptr(),
                                              self.win_returns.data_ptr(), self._stream()), "m3pc_ring_windows")
        if plan:
            assert L.cfg.plan_guidance in ("critic_lambda_guiding", "rtg_guiding", "noise_adding_lambda")
            lmbda = 0.6 if L.cfg.plan_guidance == "rtg_guiding" else L.cfg.lmbda
            guidance = L.cfg.plan_guidance
        else:
            lmbda, guidance = 0.0, "mtm_sampling"
        lead = (E,) if E > 1 else ()
        ev, sm = L._plan_device(guidance, horizon, lmbda, self.win_states.view(*lead, T, self.obs), self.win_actions.view(*lead, T, self.act),
                                self.win_rewards.view(*lead, T), self.win_returns.view(*lead, T), n_env=E)
        from .learner import PlanTicket
        pool = L.__dict__.setdefault("_tickets", {}).setdefault(E, [])
        ticket = pool.pop() if pool else PlanTicket(E, self.act, pool)
        ticket._submit(ev if eval else sm)
        return ticket


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
        a = np.asarray(action, dtype=np.float32).reshape(-1)  # the reference hands (1, act) exploration actions to env.step
        self.x = (self.A @ self.x + self.B @ a).astype(np.float32)
        self.t += 1
        reward = float(np.tanh(self.w @ self.x))
        return self.x.copy(), reward, self.t >= self.horizon, {}
