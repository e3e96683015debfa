"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the reference MTM forward.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``m3pc_b200/`` does.

What it restates: the masked-trajectory-model forward of wkh923/m3pc, as plain torch-CPU tensor
ops on an explicit ``state_dict`` (reference key names), with every function citing the
reference lines it follows.  The arithmetic of the reference lives in a third-party dependency
that is not vendored -- **PyTorch** (reference README pins 1.12.1; this image has 2.11.0):
``nn.TransformerEncoderLayer(norm_first=True, activation="gelu")``, ``nn.MultiheadAttention``,
``nn.LayerNorm(eps=1e-5)``, ``nn.Linear``, ``nn.GELU`` (erf form), ``torch.distributions.Normal``.
Their published algorithms are restated below op by op.

Parity pinning: the reference's own tests hold NO golden vector for this path,
so the oracle is pinned against outputs of the reference itself, run in the dev container by
``tests/golden/gen_golden.py`` and committed as ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks the oracle against them on CPU).

All functions are dtype-generic: pass float32 tensors for the reference's arithmetic, float64 for a
high-precision ground truth.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Mapping, Tuple

import numpy as np
import torch
import torch.nn.functional as F

KEYS = ("states", "actions", "rewards", "returns")  # call order, learner.py:361-366


# --------------------------------------------------------------------------------------
# tokenizers  (research/omtm/tokenizers/continuous.py)
# --------------------------------------------------------------------------------------
def tokenizer_encode(x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, normalize: bool) -> torch.Tensor:
    """continuous.py:68-79 -- (B,T,d) -> (B,T,1,d); ``(x-mean)/std`` unless ``normalize`` is False (actions)."""
    assert x.dim() == 3
    if normalize:
        x = (x - mean) / std
    return x.unsqueeze(2).to(mean.dtype)


def tokenizer_decode(y: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, normalize: bool) -> torch.Tensor:
    """continuous.py:81-94 -- (B,T,1,d) -> (B,T,d); ``y*std+mean``; un-normalised modalities pass through."""
    if normalize:
        assert y.dim() == 4 and y.size(2) == 1
        return y.squeeze(2) * std + mean
    return y


def encode_all(traj: Mapping[str, torch.Tensor], stats: Mapping[str, Mapping[str, torch.Tensor]]) -> "OrderedDict[str, torch.Tensor]":
    """TokenizerManager.encode, base.py:69-83; actions are not normalised (continuous.py:57-61)."""
    out = OrderedDict()
    for k, v in traj.items():
        out[k] = tokenizer_encode(v, stats[k]["mean"], stats[k]["std"], normalize=(k != "actions"))
        assert out[k].dim() == 4
    return out


# --------------------------------------------------------------------------------------
# masks -> index tables  (mtm_model.py:534-544, 559-591)
# --------------------------------------------------------------------------------------
def index_mask(mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """omtm._index, mtm_model.py:534-544: ids of kept tokens, the restore permutation, keep_len."""
    assert mask.dim() == 1
    ids = (mask == 1).nonzero(as_tuple=True)[0]
    zero_ids = (mask == 0).nonzero(as_tuple=True)[0]
    ids_restore = torch.argsort(torch.hstack((ids, zero_ids)))
    return ids, ids_restore, int(len(ids))


# --------------------------------------------------------------------------------------
# transformer pieces (PyTorch's published algorithms, restated)
# --------------------------------------------------------------------------------------
def layer_norm_plain(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm's published algorithm: over the last dim, biased variance, eps inside the sqrt."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf_plain(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default (approximate='none'): 0.5 x (1 + erf(x / sqrt 2))."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """What nn.LayerNorm calls (== layer_norm_plain, tests/test_oracle_golden.py checks it).  The library function is used
    so that the oracle, when timed as the CPU baseline, costs what the reference's own modules cost."""
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """What nn.GELU() calls (== gelu_erf_plain)."""
    return F.gelu(x)


def multi_head_attention(y: torch.Tensor, sd: Mapping[str, torch.Tensor], p: str, n_head: int) -> torch.Tensor:
    """nn.MultiheadAttention(batch_first=True), self-attention, no mask, eval mode.
    qkv = y W_in^T + b_in; per head softmax(q k^T / sqrt(d_h)) v; out projection."""
    b, s, d = y.shape
    dh = d // n_head
    qkv = F.linear(y, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"])
    q, k, v = qkv.split(d, dim=-1)
    q = q.reshape(b, s, n_head, dh).transpose(1, 2)
    k = k.reshape(b, s, n_head, dh).transpose(1, 2)
    v = v.reshape(b, s, n_head, dh).transpose(1, 2)
    # softmax(q k^T / sqrt(d_h)) v -- the library call nn.MultiheadAttention's fast path uses
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, s, d)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def encoder_layer(x: torch.Tensor, sd: Mapping[str, torch.Tensor], p: str, n_head: int) -> torch.Tensor:
    """nn.TransformerEncoderLayer(norm_first=True, activation='gelu'), eval mode (dropout off):
    x += MHA(LN1(x)); x += W2 gelu(W1 LN2(x) + b1) + b2.   Built at mtm_model.py:379-409."""
    x = x + multi_head_attention(layer_norm(x, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"]), sd, p + ".self_attn", n_head)
    h = gelu_erf(F.linear(layer_norm(x, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"]), sd[p + ".linear1.weight"], sd[p + ".linear1.bias"]))
    return x + F.linear(h, sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])


def transformer(x: torch.Tensor, sd: Mapping[str, torch.Tensor], p: str, n_layer: int, n_head: int, stages=None) -> torch.Tensor:
    """nn.TransformerEncoder(layers, norm=LayerNorm): the stack followed by the final norm.
    ``stages`` (test hook): receives the residual stream before the final norm under ``p + "_x"``."""
    for i in range(n_layer):
        x = encoder_layer(x, sd, f"{p}.layers.{i}", n_head)
    if stages is not None:
        stages[p + "_x"] = x
    return layer_norm(x, sd[p + ".norm.weight"], sd[p + ".norm.bias"])


# --------------------------------------------------------------------------------------
# the model forward  (mtm_model.py:593-607 and callees)
# --------------------------------------------------------------------------------------
def mtm_forward(
    sd: Mapping[str, torch.Tensor],
    tokens: Mapping[str, torch.Tensor],
    masks: Mapping[str, torch.Tensor],
    n_head: int,
    n_enc_layer: int,
    n_dec_layer: int,
    return_stages: bool = False,
) -> Dict[str, object]:
    """omtm.forward(trajectories, masks), P == 1 tokens per time step, continuous modalities.

    ``tokens[k]``: (B,T,1,d_k) already tokenised; ``masks[k]``: (T,) of {0,1}.
    Returns states/rewards/returns as (B,T,1,d) tensors and the action head as
    ``{"mu": ..., "std": ...}`` (the parameters of the reference's SquashedNormal,
    mtm_model.py:313-321: mean = tanh(mu), sample = tanh(mu + std * eps)).
    """
    keys = list(tokens.keys())
    pos = sd["pos_embed"]  # (1,T,1,D)
    stages: Dict[str, object] = {}

    # trajectory_encoding, mtm_model.py:546-557 (all T tokens are embedded, masked ones dropped after)
    emb = OrderedDict()
    for k in keys:
        x = tokens[k]
        e = F.linear(x, sd[f"encoder_embed_dict.{k}.weight"], sd[f"encoder_embed_dict.{k}.bias"])
        e = e + sd[f"encoder_per_dim_encoding.{k}"] + pos[:, : x.shape[1], :, :]
        b, t, p_, c = e.shape
        emb[k] = e.reshape(b, t * p_, c)

    # forward_encoder, mtm_model.py:619-644
    feats, restore, keep = [], {}, {}
    for k in keys:
        ids, restore[k], keep[k] = index_mask(masks[k].reshape(-1))
        feats.append(emb[k][:, ids])
    x = torch.cat(feats, dim=1)
    stages["enc_in"] = x
    x = transformer(x, sd, "encoder", n_enc_layer, n_head, stages)
    stages["enc_out"] = x
    enc = OrderedDict()
    idx = 0
    for k in keys:
        enc[k] = x[:, idx : idx + keep[k]]
        idx += keep[k]

    # forward_decoder, mtm_model.py:663-716 (+ _decoder_trajectory_encoding :646-661)
    dec_in = []
    for k in keys:
        b = enc[k].shape[0]
        n_mask = restore[k].shape[0] - keep[k]
        x_ = torch.cat([enc[k], sd[f"mask_token_dict.{k}"].repeat(b, n_mask, 1)], dim=1)
        x_ = torch.gather(x_, 1, restore[k][None, :, None].repeat(b, 1, x_.shape[-1]))
        e = F.linear(x_, sd[f"decoder_embed_dict.{k}.weight"], sd[f"decoder_embed_dict.{k}.bias"])
        t = e.shape[1]
        e = e + sd[f"decoder_per_dim_encoding.{k}"][:, :, 0, :] + pos[:, :t, 0, :]
        dec_in.append(e)
    x = torch.cat(dec_in, dim=1)
    stages["dec_in"] = x
    x = transformer(x, sd, "decoder", n_dec_layer, n_head, stages)
    stages["dec_out"] = x

    out: Dict[str, object] = {}
    p0 = 0
    for k, e in zip(keys, dec_in):
        t = e.shape[1]
        seg = x[:, p0 : p0 + t, :].reshape(x.shape[0], t, 1, x.shape[-1])
        p0 += t
        if k == "actions":
            # DiagGaussianActor.forward, mtm_model.py:313-321, log_std_bounds = [-5, 2]
            mu = F.linear(seg, sd["output_head_dict.actions.mu.weight"], sd["output_head_dict.actions.mu.bias"])
            ls = torch.tanh(F.linear(seg, sd["output_head_dict.actions.log_std.weight"], sd["output_head_dict.actions.log_std.bias"]))
            ls = -5.0 + 0.5 * (2.0 - (-5.0)) * (ls + 1.0)
            out[k] = {"mu": mu, "std": ls.exp()}
        else:
            # nn.Sequential(LayerNorm, Linear, GELU, Linear), mtm_model.py:428-433
            h = layer_norm(seg, sd[f"output_head_dict.{k}.0.weight"], sd[f"output_head_dict.{k}.0.bias"])
            h = gelu_erf(F.linear(h, sd[f"output_head_dict.{k}.1.weight"], sd[f"output_head_dict.{k}.1.bias"]))
            out[k] = F.linear(h, sd[f"output_head_dict.{k}.3.weight"], sd[f"output_head_dict.{k}.3.bias"])
    if return_stages:
        out["_stages"] = stages
    return out


def decode_all(pred: Mapping[str, object], stats: Mapping[str, Mapping[str, torch.Tensor]]) -> Dict[str, object]:
    """TokenizerManager.decode, base.py:85-99: de-normalise states/rewards/returns, pass the action dist through."""
    out: Dict[str, object] = {}
    for k, v in pred.items():
        if k.startswith("_"):
            continue
        if k == "actions":
            out[k] = v
        else:
            out[k] = tokenizer_decode(v, stats[k]["mean"], stats[k]["std"], normalize=True)
    return out


# --------------------------------------------------------------------------------------
# TwinQ critic  (research/finetune_omtm/model.py:72-104, 146-171)
# --------------------------------------------------------------------------------------
def twinq(qsd: Mapping[str, torch.Tensor], obs_mean: torch.Tensor, obs_std: torch.Tensor, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
    """TwinQ.forward: min(q1, q2) of two ReLU MLPs [obs+act, 256, 256, 1] on ((s - mu_o)/sigma_o, a)."""
    sa = torch.cat([(state - obs_mean) / obs_std, action], 1)
    qs = []
    for q in ("q1", "q2"):
        h = torch.relu(F.linear(sa, qsd[f"{q}.net.0.weight"], qsd[f"{q}.net.0.bias"]))
        h = torch.relu(F.linear(h, qsd[f"{q}.net.2.weight"], qsd[f"{q}.net.2.bias"]))
        qs.append(F.linear(h, qsd[f"{q}.net.4.weight"], qsd[f"{q}.net.4.bias"]).squeeze(-1))
    return torch.min(qs[0], qs[1])


def to_torch(d: Mapping[str, np.ndarray], dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, torch.as_tensor(np.asarray(v)).to(dtype)) for k, v in d.items())


def stats_to_torch(stats: Mapping[str, Mapping[str, np.ndarray]], dtype=torch.float32):
    return {k: {n: torch.as_tensor(np.asarray(a)).to(dtype) for n, a in s.items()} for k, s in stats.items()}
