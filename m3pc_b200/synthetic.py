"""Deterministic synthetic inputs for the M^3PC planning hot path.

Everything here is drawn from ``numpy.random.RandomState`` (bit-stable across numpy
versions and machines), never from torch's generator, so that the golden fixtures in
``tests/golden/`` (produced by running the *reference* in the dev container), the parity
tests on the GPU box and ``bench.py`` all see identical weights, tokenizer statistics,
history windows and injected noise.

Shapes follow the reference:
  * model parameters: ``research/omtm/models/mtm_model.py:324-437`` (state_dict key names
    are the reference's);
  * tokenizer statistics: ``research/omtm/tokenizers/continuous.py:32-62``;
  * TwinQ critic: ``research/finetune_omtm/model.py:146-171``;
  * history dict: ``research/finetune_omtm/learner.py:342-366``.

The reference zero-initialises mask tokens, per-dim encodings and (for the actor head)
biases (``mtm_model.py:364,372-377,304-309``); every such tensor is re-drawn ~N(0, 0.2^2)
here, otherwise the mask-token / per-dim paths would be vacuous in a parity check.
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

MODALITIES = ("states", "actions", "rewards", "returns")

#: D4RL shapes used by BASELINE.json's configs: name -> (obs_dim, act_dim)
ENV_SHAPES = {
    "hopper": (11, 3),
    "walker2d": (17, 6),
    "halfcheetah": (17, 6),
}


@dataclasses.dataclass(frozen=True)
class ModelShape:
    """Static shape description of one MTM instance (mirrors ``omtmConfig`` + data_shapes)."""

    obs_dim: int
    act_dim: int
    n_embd: int = 512
    n_head: int = 4
    n_enc_layer: int = 2
    n_dec_layer: int = 1
    traj_length: int = 8

    @property
    def feature_dims(self) -> Dict[str, int]:
        return {"states": self.obs_dim, "actions": self.act_dim, "rewards": 1, "returns": 1}

    @property
    def data_shapes(self) -> "OrderedDict[str, Tuple[int, int]]":
        # key order is the reference's call order (learner.py:361-366)
        return OrderedDict((k, (1, d)) for k, d in self.feature_dims.items())


def shipped_shape(env: str = "hopper") -> ModelShape:
    """The shipped planning config: finetune_omtm/config.yaml:27-39 (D=512, 4 heads, 2+1 layers, T=8)."""
    o, a = ENV_SHAPES[env]
    return ModelShape(obs_dim=o, act_dim=a)


def scaled_shape(env: str = "hopper") -> ModelShape:
    """BASELINE.json config 5: 2x embed dim, 2x layers, 2x horizon (head_dim kept at 128)."""
    o, a = ENV_SHAPES[env]
    return ModelShape(obs_dim=o, act_dim=a, n_embd=1024, n_head=8, n_enc_layer=4, n_dec_layer=2, traj_length=16)


def _linear(rs: np.random.RandomState, out_f: int, in_f: int, prefix: str, sd: dict, bias_std: float = 0.2):
    bound = 1.0 / np.sqrt(in_f)
    sd[prefix + ".weight"] = rs.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)
    sd[prefix + ".bias"] = (rs.randn(out_f) * bias_std).astype(np.float32)


def _layernorm(rs: np.random.RandomState, d: int, prefix: str, sd: dict):
    sd[prefix + ".weight"] = rs.uniform(0.8, 1.2, size=(d,)).astype(np.float32)
    sd[prefix + ".bias"] = (rs.randn(d) * 0.1).astype(np.float32)


def _transformer(rs: np.random.RandomState, d: int, n_layer: int, prefix: str, sd: dict):
    for i in range(n_layer):
        p = f"{prefix}.layers.{i}"
        bound = np.sqrt(6.0 / (d + 3 * d))  # xavier_uniform on in_proj_weight, as nn.MultiheadAttention
        sd[p + ".self_attn.in_proj_weight"] = rs.uniform(-bound, bound, size=(3 * d, d)).astype(np.float32)
        sd[p + ".self_attn.in_proj_bias"] = (rs.randn(3 * d) * 0.2).astype(np.float32)
        _linear(rs, d, d, p + ".self_attn.out_proj", sd)
        _linear(rs, 4 * d, d, p + ".linear1", sd)
        _linear(rs, d, 4 * d, p + ".linear2", sd)
        _layernorm(rs, d, p + ".norm1", sd)
        _layernorm(rs, d, p + ".norm2", sd)
    _layernorm(rs, d, prefix + ".norm", sd)


def sincos_pos_embed(n_embd: int, traj_length: int) -> np.ndarray:
    """The reference's fixed buffer ``pos_embed`` (1,T,1,D): mtm_model.py:38-58 and :435-437 (note the /2)."""
    omega = np.arange(n_embd // 2, dtype=np.float32)
    omega /= n_embd / 2.0
    omega = 1.0 / 10000 ** omega
    pos = np.arange(traj_length, dtype=np.float32).reshape(-1)
    out = np.einsum("m,d->md", pos, omega)
    emb = np.concatenate([np.sin(out), np.cos(out)], axis=1)
    # same op order as the reference: float32 tensor, then / 2.0
    return (emb.astype(np.float32)[None, :, None, :] / np.float32(2.0)).astype(np.float32)


def make_state_dict(shape: ModelShape, seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """A full ``omtm.state_dict()`` (reference key names) with every tensor non-trivial."""
    rs = np.random.RandomState(seed)
    d = shape.n_embd
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for k, f in shape.feature_dims.items():
        sd[f"mask_token_dict.{k}"] = (rs.randn(1, 1, d) * 0.2).astype(np.float32)
    for k, f in shape.feature_dims.items():
        sd[f"encoder_per_dim_encoding.{k}"] = (rs.randn(1, 1, 1, d) * 0.2).astype(np.float32)
    for k, f in shape.feature_dims.items():
        sd[f"decoder_per_dim_encoding.{k}"] = (rs.randn(1, 1, 1, d) * 0.2).astype(np.float32)
    for k, f in shape.feature_dims.items():
        _linear(rs, d, f, f"encoder_embed_dict.{k}", sd)
    for k, f in shape.feature_dims.items():
        _linear(rs, d, d, f"decoder_embed_dict.{k}", sd)
    _transformer(rs, d, shape.n_enc_layer, "encoder", sd)
    _transformer(rs, d, shape.n_dec_layer, "decoder", sd)
    for k, f in shape.feature_dims.items():
        if k == "actions":
            _linear(rs, f, d, "output_head_dict.actions.mu", sd)
            _linear(rs, f, d, "output_head_dict.actions.log_std", sd)
        else:
            _layernorm(rs, d, f"output_head_dict.{k}.0", sd)
            _linear(rs, d, d, f"output_head_dict.{k}.1", sd)
            _linear(rs, f, d, f"output_head_dict.{k}.3", sd)
    sd["pos_embed"] = sincos_pos_embed(d, shape.traj_length)
    return sd


def make_tokenizer_stats(shape: ModelShape, seed: int = 1) -> Dict[str, Dict[str, np.ndarray]]:
    """Per-modality mean/std/min/max: mean~N(0,1), std~U(0.5,1.5), min/max = mean -/+ 3 std."""
    rs = np.random.RandomState(seed)
    stats = {}
    for k, f in shape.feature_dims.items():
        mean = rs.randn(f).astype(np.float32)
        std = rs.uniform(0.5, 1.5, size=(f,)).astype(np.float32)
        stats[k] = {"mean": mean, "std": std, "min": mean - 3 * std, "max": mean + 3 * std}
    return stats


def make_critic_state_dict(shape: ModelShape, seed: int = 2, hidden: int = 256) -> "OrderedDict[str, np.ndarray]":
    """``TwinQ.state_dict()``: q{1,2}.net.{0,2,4}.{weight,bias} (finetune_omtm/model.py:72-104,146-160)."""
    rs = np.random.RandomState(seed)
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    dims = [shape.obs_dim + shape.act_dim, hidden, hidden, 1]
    for q in ("q1", "q2"):
        for li, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
            _linear(rs, o, i, f"{q}.net.{2 * li}", sd, bias_std=0.1)
    return sd


def make_obs_norm(shape: ModelShape, seed: int = 3) -> Tuple[np.ndarray, np.ndarray]:
    rs = np.random.RandomState(seed)
    return rs.randn(shape.obs_dim).astype(np.float32) * 0.5, rs.uniform(0.5, 1.5, size=(shape.obs_dim,)).astype(np.float32)


def make_history(shape: ModelShape, seed: int = 4, length: int = 1000, path_length: int = 50) -> Dict[str, np.ndarray]:
    """The ``sequence_history`` dict the callers hand to ``action_sample`` (replay_buffer.py:204-212)."""
    rs = np.random.RandomState(seed)
    return {
        "observations": rs.randn(length, shape.obs_dim).astype(np.float32),
        "actions": rs.uniform(-1, 1, size=(length, shape.act_dim)).astype(np.float32),
        "rewards": rs.randn(length, 1).astype(np.float32),
        "values": np.zeros((length, 1), dtype=np.float32),
        "path_length": int(path_length),
    }


def make_noise(n_cand: int, traj_length: int, act_dim: int, seed: int = 7) -> Tuple[np.ndarray, np.ndarray]:
    """Injected noise: eps ~ N(0,1) of shape (N,1,T,1,A) (what ``SquashedNormal.sample((N,))`` consumes,
    learner.py:285-287) and q ~ Exp(1) of shape (N,) (what ``torch.multinomial`` consumes, learner.py:324)."""
    rs = np.random.RandomState(seed)
    eps = rs.randn(n_cand, 1, traj_length, 1, act_dim).astype(np.float32)
    q = rs.exponential(1.0, size=(n_cand,)).astype(np.float32)
    return eps, q


def make_trajectories(shape: ModelShape, batch: int, seed: int = 5) -> Dict[str, np.ndarray]:
    """Raw (un-tokenised) trajectories (B,T,d) for direct ``omtm.forward`` parity checks."""
    rs = np.random.RandomState(seed)
    t = shape.traj_length
    return {
        "states": rs.randn(batch, t, shape.obs_dim).astype(np.float32),
        "actions": rs.uniform(-1, 1, size=(batch, t, shape.act_dim)).astype(np.float32),
        "rewards": rs.randn(batch, t, 1).astype(np.float32),
        "returns": rs.randn(batch, t, 1).astype(np.float32),
    }
