"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference M^3PC planners.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Restates, loop for loop (no closed forms, no dead-row elimination -- this is the checker and
the CPU baseline, so it does the work the reference does):
  * inference masks      research/finetune_omtm/masks.py:7-44, research/zeroshot_omtm/masks.py:30-91
  * window builder       research/finetune_omtm/learner.py:329-417 (``Learner.action_sample``)
  * rtg_guiding          research/finetune_omtm/learner.py:271-327
  * critic_lambda_guiding research/finetune_omtm/learner.py:211-268
  * noise_adding_lambda  research/finetune_omtm/learner.py:142-208
  * mtm_sampling         research/finetune_omtm/learner.py:103-115
  * zero-shot planners   research/zeroshot_omtm/learner.py:60-149 (id), :151-261 (piid)

Randomness is *injected*: ``eps`` replaces the N(0,1) draws of ``SquashedNormal.sample`` /
``torch.randn`` and ``q`` the Exp(1) draws inside ``torch.multinomial`` (SURVEY.md section 8c:
``multinomial(p,1) == argmax(p / q)``).  Pinned against the live reference by
``tests/golden/gen_golden.py`` -> ``tests/golden/*.npz`` -> ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict
from typing import Dict, Mapping, Optional

import numpy as np
import torch

from . import mtm_oracle as mo


# --------------------------------------------------------------------------------------
# masks (numpy float64 vectors of {0,1}, exactly the reference's values)
# --------------------------------------------------------------------------------------
def _pack(s, a, rw, rt) -> "OrderedDict[str, np.ndarray]":
    return OrderedDict([("states", s), ("actions", a), ("rewards", rw), ("returns", rt)])


def create_rcbc_mask(T: int, idx: int):
    """finetune_omtm/masks.py:7-27."""
    s = np.zeros(T); s[: idx + 1] = 1
    rt = np.ones(T)
    a = np.zeros(T)
    if idx > 0:
        a[:idx] = 1
    return _pack(s, a, np.zeros(T), rt)


def create_fd_mask(T: int, idx: int):
    """finetune_omtm/masks.py:30-44."""
    s = np.zeros(T); s[: idx + 1] = 1
    return _pack(s, np.ones(T), np.zeros(T), np.zeros(T))


def create_fid_mask(T: int, idx: int):
    """zeroshot_omtm/masks.py:30-47."""
    a = np.zeros(T)
    if idx > 0:
        a[:idx] = 1
    return _pack(np.ones(T), a, np.zeros(T), np.zeros(T))


def create_gid_mask(T: int, idx: int):
    """zeroshot_omtm/masks.py:50-69 (create_pi_mask :72-91 is the same function)."""
    s = np.ones(T)
    if idx > 0:
        s[idx + 1 : -1] = 0
    a = np.zeros(T)
    if idx > 0:
        a[:idx] = 1
    return _pack(s, a, np.zeros(T), np.zeros(T))


create_pi_mask = create_gid_mask


def _tmask(m, dtype=torch.float64):
    return OrderedDict((k, torch.from_numpy(v).to(dtype)) for k, v in m.items())


# --------------------------------------------------------------------------------------
# the Learner state the planners need
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class PlannerOracle:
    sd: Mapping[str, torch.Tensor]            # omtm state_dict
    stats: Mapping[str, Mapping[str, torch.Tensor]]  # tokenizer mean/std/min/max per modality
    n_head: int
    n_enc_layer: int
    n_dec_layer: int
    traj_length: int
    action_samples: int
    horizon: int = 4
    discount: float = 0.99
    temperature: float = 0.01
    lmbda: float = 0.6
    plan_guidance: str = "rtg_guiding"
    critic_sd: Optional[Mapping[str, torch.Tensor]] = None
    obs_mean: Optional[torch.Tensor] = None
    obs_std: Optional[torch.Tensor] = None
    #: test hook: (N,h,A) candidate actions used INSTEAD of the ones derived from ``eps`` -- lets a test score the candidates
    #: a device run drew from its own Philox stream (the reference has no such hook; everything downstream is unchanged)
    cand_override: Optional[torch.Tensor] = None

    # -- model call: tokenizer_manager.decode(self.mtm(tokenizer_manager.encode(traj), mask)) --
    def _model(self, traj, mask):
        enc = mo.encode_all(traj, self.stats)
        pred = mo.mtm_forward(self.sd, enc, _tmask(mask), self.n_head, self.n_enc_layer, self.n_dec_layer)
        return mo.decode_all(pred, self.stats)

    @property
    def dtype(self):
        return self.sd["pos_embed"].dtype

    # -- learner.py:103-115 --
    def mtm_sampling(self, traj, h, eps1=None):
        T = self.traj_length
        dist = self._model(traj, create_rcbc_mask(T, T - h))["actions"]
        mu, std = dist["mu"], dist["std"]
        e = torch.zeros_like(mu) if eps1 is None else eps1.reshape(mu.shape)
        sample_action = torch.tanh(mu + std * e)[0, T - h]
        eval_action = torch.tanh(mu)[0, T - h]
        return sample_action, eval_action, {}

    # -- shared tail: learner.py:318-325 --
    @staticmethod
    def _select(expect_return, sample_actions, temperature, q):
        expect_return = expect_return - torch.max(expect_return)
        score = expect_return * temperature
        p = torch.exp(score) / torch.exp(score).sum()
        eval_action = (sample_actions[:, 0] * p[:, None]).sum(dim=0) / p.sum()
        if q is None:
            sample_idx = torch.argmax(p).reshape(1)
        else:
            sample_idx = torch.argmax(p / q).reshape(1)  # == torch.multinomial(p, 1) given its Exp(1) draws
        sample_action = sample_actions[sample_idx, 0]
        return sample_action, eval_action, p, sample_idx

    def _candidates_from_dist(self, dist, h, eps):
        """learner.py:285-287: action_dist.sample((N,))[:, 0, T-h:, 0, :]; eps has shape (N,1,T,1,A)."""
        T = self.traj_length
        mu, std = dist["mu"], dist["std"]  # (1,T,1,A)
        return torch.tanh(mu[None] + std[None] * eps)[:, 0, T - h :, 0, :]

    def _guided(self, traj, h, lmbda, eps, q, value_kind, cand_kind):
        T, N = self.traj_length, self.action_samples
        batch = {k: v.repeat(N, 1, 1) for k, v in traj.items()}
        dist = self._model(traj, create_rcbc_mask(T, T - h))["actions"]
        if self.cand_override is not None:
            sample_actions = None
        elif cand_kind == "dist":
            sample_actions = self._candidates_from_dist(dist, h, eps)
        else:  # noise_adding_lambda, learner.py:156-167: eps has shape (N,h,A)
            mean = torch.tanh(dist["mu"])[0, T - h :, 0, :]
            sample_actions = torch.clamp(mean + eps * 0.09, -0.99999, 0.99999)
        if self.cand_override is not None:
            sample_actions = self.cand_override.to(self.dtype).reshape(N, h, -1)
        batch["actions"] = batch["actions"].clone()
        batch["actions"][:, T - h :, :] = sample_actions
        dec = self._model(batch, create_fd_mask(T, T - h))
        future_states = dec["states"][:, T - h :, :]
        future_rewards = dec["rewards"][:, T - h :, :]
        expect_return = torch.zeros((N,), dtype=self.dtype)
        for t in range(h):
            values = torch.zeros((N, t + 1), dtype=self.dtype)
            discounts = torch.cumprod(self.discount * torch.ones((t + 1,), dtype=self.dtype), dim=0)
            if value_kind == "rtg":
                values[:, t] = dec["returns"][:, T - h + t, 0] * 1000
                if t > 0:
                    values[:, :t] = future_rewards[:, :t, 0]
            else:
                if t > 0:
                    values[:, :t] = future_rewards[:, :t, 0]
                values[:, t] = mo.twinq(self.critic_sd, self.obs_mean, self.obs_std, future_states[:, t], sample_actions[:, t])
            values = values * discounts[None, :]
            if t < h - 1:
                expect_return = expect_return + values.sum(dim=-1) * (1 - lmbda) * (lmbda ** t)
            else:
                expect_return = expect_return + values.sum(dim=-1) * (lmbda ** t)
        sample_action, eval_action, p, sample_idx = self._select(expect_return, sample_actions, self.temperature, q)
        dbg = {
            "candidates": sample_actions,
            "expect_return": expect_return,
            "p": p,
            "sample_idx": sample_idx,
            "argmax": torch.argmax(expect_return),
            "act_mu": dist["mu"],
            "act_std": dist["std"],
            "future_states": future_states,
            "future_rewards": future_rewards,
            "future_returns": dec["returns"][:, T - h :, :],
        }
        return sample_action, eval_action, dbg

    def rtg_guiding(self, traj, h, eps, q, lmbda=0.6):
        """learner.py:271-327 (lmbda is NOT forwarded by action_sample: fixed 0.6, :405-407)."""
        return self._guided(traj, h, lmbda, eps, q, "rtg", "dist")

    def critic_lambda_guiding(self, traj, h, lmbda, eps, q):
        """learner.py:211-268."""
        return self._guided(traj, h, lmbda, eps, q, "critic", "dist")

    def noise_adding_lambda(self, traj, h, lmbda, eps, q):
        """learner.py:142-208."""
        return self._guided(traj, h, lmbda, eps, q, "critic", "noise")

    # -- learner.py:342-385: the host-side window builder --
    def build_window(self, hist, percentage=1.0, rtg=None, future_obs=False):
        T = self.traj_length
        horizon = self.horizon
        end_idx = int(hist["path_length"])
        if end_idx + horizon < T:
            horizon = T - end_idx
        obs_dim = hist["observations"].shape[-1]
        act_dim = hist["actions"].shape[-1]
        zero = {
            "observations": np.zeros((1, T, obs_dim)),
            "actions": np.zeros((1, T, act_dim)),
            "rewards": np.zeros((1, T, 1)),
            "values": np.zeros((1, T, 1)),
        }
        hl = T - horizon + 1
        for k in zero:
            zero[k][0, :hl] = hist[k][end_idx - hl + 1 : end_idx + 1]
        if future_obs:  # zeroshot_omtm/learner.py:75-79, 97-106
            smart_T = T
            if end_idx + horizon > 1000:
                smart_T = smart_T - (end_idx + horizon - 1000)
            zero["observations"][0, :smart_T] = hist["observations"][end_idx - hl + 1 : end_idx - hl + 1 + T]
        traj = OrderedDict()
        for k, v in zero.items():
            name = "states" if k == "observations" else "returns" if k == "values" else k
            traj[name] = torch.tensor(v, dtype=torch.float32).to(self.dtype)
        if rtg is not None:
            rtg_v = float(rtg)
        else:
            rmax = self.stats["returns"]["max"]
            rmin = self.stats["returns"]["min"]
            rtg_v = float(rmin + (rmax - rmin) * percentage)
        # float64 until the tokenizer casts it (continuous.py:79)
        traj["returns"] = torch.from_numpy(rtg_v * np.ones((1, T, 1))).to(torch.float64 if self.dtype == torch.float32 else self.dtype)
        return traj, horizon

    def action_sample(self, hist, percentage=1.0, plan=True, eval=False, rtg=None, eps=None, q=None):
        """learner.py:329-417.  Returns (action, debug dict)."""
        if eval:
            assert rtg is not None
        traj, h = self.build_window(hist, percentage, rtg)
        if plan:
            assert self.plan_guidance in ["critic_lambda_guiding", "rtg_guiding", "noise_adding_lambda"]
            if self.plan_guidance == "critic_lambda_guiding":
                s, e, dbg = self.critic_lambda_guiding(traj, h, self.lmbda, eps, q)
            elif self.plan_guidance == "noise_adding_lambda":
                s, e, dbg = self.noise_adding_lambda(traj, h, self.lmbda, eps, q)
            else:
                s, e, dbg = self.rtg_guiding(traj, h, eps, q)
        else:
            s, e, dbg = self.mtm_sampling(traj, h, eps)
        dbg["horizon"] = h
        dbg["sample_action"], dbg["eval_action"] = s, e
        return (e if eval else s), dbg

    # -- zeroshot_omtm/learner.py:60-149 --
    def action_id_sample(self, hist, percentage=1.0, eval=False, rtg=None, eps1=None):
        if eval:
            assert rtg is not None
        T = self.traj_length
        traj, h = self.build_window(hist, percentage, rtg, future_obs=True)
        dist = self._model(traj, create_gid_mask(T, T - h))["actions"]
        mu, std = dist["mu"], dist["std"]
        e = torch.zeros_like(mu) if eps1 is None else eps1.reshape(mu.shape)
        sample_action = torch.tanh(mu + std * e)[0, T - h]
        eval_action = torch.tanh(mu)[0, T - h]
        return (eval_action if eval else sample_action), {"horizon": h, "sample_action": sample_action, "eval_action": eval_action}

    # -- zeroshot_omtm/learner.py:151-261 --
    def action_piid_sample(self, hist, percentage=1.0, eval=False, rtg=None, eps1=None):
        if eval:
            assert rtg is not None
        T = self.traj_length
        traj, h = self.build_window(hist, percentage, rtg, future_obs=True)
        state_inf = self._model(traj, create_pi_mask(T, T - h))["states"]
        traj["states"] = traj["states"].clone()
        traj["states"][:, T - h + 2 : -1, :] = state_inf[:, T - h + 2 : -1, :]
        traj["states"][:, : T - h + 1, :] = state_inf[:, : T - h + 1, :]
        dist = self._model(traj, create_fid_mask(T, T - h))["actions"]
        mu, std = dist["mu"], dist["std"]
        e = torch.zeros_like(mu) if eps1 is None else eps1.reshape(mu.shape)
        sample_action = torch.tanh(mu + std * e)[0, T - h]
        eval_action = torch.tanh(mu)[0, T - h]
        dbg = {"horizon": h, "sample_action": sample_action, "eval_action": eval_action,
               "state_inference": state_inf, "states_filled": traj["states"]}
        return (eval_action if eval else sample_action), dbg


def from_synthetic(shape, sd_np, stats_np, dtype=torch.float32, critic_np=None, obs_norm=None, **kw) -> PlannerOracle:
    """Build a PlannerOracle from the numpy dicts of ``m3pc_b200.synthetic``."""
    critic = mo.to_torch(critic_np, dtype) if critic_np is not None else None
    om = torch.as_tensor(obs_norm[0]).to(dtype) if obs_norm is not None else None
    os_ = torch.as_tensor(obs_norm[1]).to(dtype) if obs_norm is not None else None
    return PlannerOracle(
        sd=mo.to_torch(sd_np, dtype), stats=mo.stats_to_torch(stats_np, dtype),
        n_head=shape.n_head, n_enc_layer=shape.n_enc_layer, n_dec_layer=shape.n_dec_layer,
        traj_length=shape.traj_length, critic_sd=critic, obs_mean=om, obs_std=os_, **kw,
    )
