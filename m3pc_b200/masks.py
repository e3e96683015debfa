"""Inference mask patterns of the M^3PC planners -- same names, signatures and values as the reference.

  create_rcbc_mask / create_fd_mask   research/finetune_omtm/masks.py:7-44
  create_fid_mask / create_gid_mask / create_pi_mask   research/zeroshot_omtm/masks.py:30-91

Each returns ``Dict[str, float64 Tensor(T,)]`` of {0,1} on ``device``, keys in the reference's order
(states, actions, rewards, returns).  ``mask_bits`` is the host-side layout table the CUDA engine consumes.
"""
from __future__ import annotations

from collections import OrderedDict
from collections.abc import Sequence
from typing import Dict

import numpy as np
import torch

KEYS = ("states", "actions", "rewards", "returns")


def _layout(kind: str, T: int, idx: int) -> np.ndarray:
    """(4, T) float64 array of {0,1}: 1 = the encoder sees the token."""
    if not 0 <= idx < T:
        raise ValueError(f"idx must be in [0, {T}), got {idx}")
    m = np.zeros((4, T))
    t = np.arange(T)
    hist_actions = (t < idx) if idx > 0 else np.zeros(T, dtype=bool)
    if kind == "rcbc":       # states <= idx, actions < idx, every return, no reward
        m[0] = t <= idx
        m[1] = hist_actions
        m[3] = 1
    elif kind == "fd":       # states <= idx, every action
        m[0] = t <= idx
        m[1] = 1
    elif kind == "fid":      # every state, actions < idx
        m[0] = 1
        m[1] = hist_actions
    elif kind in ("gid", "pi"):  # every state except the open interval (idx, T-1), actions < idx
        m[0] = 1
        if idx > 0:
            m[0, idx + 1:T - 1] = 0
        m[1] = hist_actions
    else:
        raise ValueError(f"unknown mask kind {kind!r}")
    return m


def _to_dict(m: np.ndarray, device) -> Dict[str, torch.Tensor]:
    return OrderedDict((k, torch.from_numpy(m[i].copy()).to(device)) for i, k in enumerate(KEYS))


def create_rcbc_mask(traj_length: int, device, idx: int) -> Dict[str, torch.Tensor]:
    """Return-conditioned behaviour cloning: predict the action at idx."""
    return _to_dict(_layout("rcbc", traj_length, idx), device)


def create_fd_mask(traj_length: int, device, idx: int) -> Dict[str, torch.Tensor]:
    """Forward dynamics: predict states / rewards / returns after idx given all actions."""
    return _to_dict(_layout("fd", traj_length, idx), device)


def create_fid_mask(traj_length: int, device, idx: int) -> Dict[str, torch.Tensor]:
    """Full inverse dynamics: every state visible, actions before idx."""
    return _to_dict(_layout("fid", traj_length, idx), device)


def create_gid_mask(traj_length: int, device, idx: int) -> Dict[str, torch.Tensor]:
    """Goal-conditioned inverse dynamics: history, the next state and the final (goal) state visible."""
    return _to_dict(_layout("gid", traj_length, idx), device)


def create_pi_mask(traj_length: int, device, idx: int) -> Dict[str, torch.Tensor]:
    """Path inference (identical layout to gid in the reference)."""
    return _to_dict(_layout("pi", traj_length, idx), device)


def mask_bits(kind: str, traj_length: int, idx: int) -> np.ndarray:
    """uint8 (4*T,) modality-major layout as ``m3pc_forward`` takes it."""
    return _layout(kind, traj_length, idx).astype(np.uint8).reshape(-1)


# ---------------------------------------------------------------------------------------------------------------------
# Validation-loss masks (research/finetune_omtm/masks.py:64-125): the random autoregressive pattern compute_mtm_loss draws
# (learner.py:432-439).  Host logic only.  The numpy generator is consumed in the reference's order -- one ``choice`` (only
# when several ratios are configured) and one ``shuffle`` per modality in ``data_shapes`` order, after one ``choice`` of the
# cut modality and one ``randint`` of the cut position -- so the same ``np.random.seed`` gives the same masks.
_CUT_ORDER = ("states", "returns", "actions", "rewards")


def create_full_random_mask(data_shape, traj_length: int, mask_ratios, device, rnd_state=None) -> torch.Tensor:
    """(T, P) float64 tensor with ``int(T * P * ratio)`` ones (visible tokens) at uniformly random places."""
    rng = np.random if rnd_state is None else rnd_state
    n = traj_length * int(data_shape[0])
    ratio = rng.choice(mask_ratios) if isinstance(mask_ratios, Sequence) else mask_ratios  # tuples, lists, OmegaConf lists
    visible = int(n * float(ratio))
    flat = np.zeros(n)
    flat[:visible] = 1.0
    rng.shuffle(flat)
    return torch.tensor(flat, device=device).reshape(traj_length, int(data_shape[0]))


def create_random_autoregressize_mask(data_shapes, mask_ratios, traj_length: int, device, p_weights=(0, 0, 0.7, 0.3)) -> Dict[str, torch.Tensor]:
    """Random visibility per modality, then everything after a random cut is hidden: time steps > cut for the modalities that
    come before the drawn one in (states, returns, actions, rewards), time steps >= cut for the drawn one and those after it;
    the last action is hidden when every action would otherwise be visible."""
    cut_mode = np.random.choice(_CUT_ORDER, p=p_weights)
    cut = np.random.randint(0, traj_length)
    out = OrderedDict((k, create_full_random_mask(shape, traj_length, mask_ratios, device)) for k, shape in data_shapes.items())
    first_hidden = cut + 1
    for k in _CUT_ORDER:
        if k == cut_mode:
            first_hidden = cut
        if k in out:
            out[k][first_hidden:, :] = 0
    if bool(out["actions"].eq(1).all()):
        out["actions"][-1] = 0
    return out
