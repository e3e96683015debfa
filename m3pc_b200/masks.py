"""Inference mask patterns of the M^3PC planners -- same names, signatures and values as the reference.

  create_rcbc_mask / create_fd_mask   research/finetune_omtm/masks.py:7-44
  create_fid_mask / create_gid_mask / create_pi_mask   research/zeroshot_omtm/masks.py:30-91

Each returns ``Dict[str, float64 Tensor(T,)]`` of {0,1} on ``device``, keys in the reference's order
(states, actions, rewards, returns).  ``mask_bits`` is the host-side layout table the CUDA engine consumes.
"""
from __future__ import annotations

from collections import OrderedDict
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
