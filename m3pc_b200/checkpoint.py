"""Ingestion of the reference's on-disk formats at the boundary of the planning path.

  mtm_<step>.pt            torch.save({"model": omtm.state_dict(), "optimizer": ..., "step": ...})   finetune.py:319-326,
                           read back by Learner.__init__ as torch.load(path)["model"]                learner.py:33-35
  iql_<step>.pt            torch.save(ImplicitQLearning.state_dict()): {"qf": TwinQ.state_dict(), "vf": ..., "actor": ..., optimisers}
                                                                                                     finetune.py:327, model.py:310-334
  d4rl_statistics_*.pkl    pickle of {"states" | "actions" | "rewards" | "returns" (older files: "values"): DataStatistics}
                           written by SequenceDataset.trajectory_statistics                          sequence_dataset.py:357-403

Nothing here touches the GPU: the functions return plain state dicts / tokenizer managers that ``omtm.load_state_dict``,
``TwinQ.load_state_dict`` and the ``Learner`` constructors take; the packed bf16 / fp32 device layout is built from those by
``PlanEngine.load_state_dict`` + ``m3pc_finalize_params`` on first use.
"""
from __future__ import annotations

import io
import os
import pickle
from typing import Dict, Mapping, Optional, Tuple

import numpy as np
import torch

from .tokenizers import ContinuousTokenizer, DataStatistics, TokenizerManager

MODALITIES = ("states", "actions", "rewards", "returns")


class CheckpointError(ValueError):
    """The file is readable but does not hold what the planning path needs."""


def _torch_load(path: str, trust_pickle: bool = False):
    """``torch.load`` restricted to tensors and plain containers (``weights_only=True``).  A file that needs anything else --
    i.e. one whose pickle stream names arbitrary Python globals, which is how a crafted checkpoint executes code -- is
    REFUSED with a ``CheckpointError``; there is no silent retry.  ``trust_pickle=True`` is the explicit opt-in for files the
    caller produced themselves (e.g. an optimiser state holding custom objects)."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if trust_pickle:
        return torch.load(path, map_location="cpu", weights_only=False)
    try:
        return torch.load(path, map_location="cpu", weights_only=True)
    except pickle.UnpicklingError as exc:
        raise CheckpointError(f"{path}: not loadable as tensors + plain containers ({str(exc).splitlines()[0][:200]}). The file "
                              "references Python objects outside torch's safe allow-list; loading it would run its pickle "
                              "stream unrestricted. If you wrote this file yourself, pass trust_pickle=True.") from exc


def _strip_prefix(sd: Mapping[str, torch.Tensor], prefix: str) -> Dict[str, torch.Tensor]:
    if sd and all(k.startswith(prefix) for k in sd):
        return {k[len(prefix):]: v for k, v in sd.items()}
    return dict(sd)


def expected_mtm_keys(data_shapes: Mapping[str, Tuple[int, int]], n_embd: int, n_enc_layer: int, n_dec_layer: int, traj_length: int
                      ) -> Dict[str, Tuple[int, ...]]:
    """Every key of ``omtm.state_dict()`` with its shape (mtm_model.py:355-433)."""
    D, F = n_embd, 4 * n_embd
    out: Dict[str, Tuple[int, ...]] = {"pos_embed": (1, traj_length, 1, D)}
    for k, (P, d) in data_shapes.items():
        out[f"encoder_embed_dict.{k}.weight"] = (D, d)
        out[f"encoder_embed_dict.{k}.bias"] = (D,)
        out[f"decoder_embed_dict.{k}.weight"] = (D, D)
        out[f"decoder_embed_dict.{k}.bias"] = (D,)
        out[f"mask_token_dict.{k}"] = (1, 1, D)
        out[f"encoder_per_dim_encoding.{k}"] = (1, 1, P, D)
        out[f"decoder_per_dim_encoding.{k}"] = (1, 1, P, D)
        if k == "actions":
            for head in ("mu", "log_std"):
                out[f"output_head_dict.{k}.{head}.weight"] = (d, D)
                out[f"output_head_dict.{k}.{head}.bias"] = (d,)
        else:
            out[f"output_head_dict.{k}.0.weight"] = (D,)
            out[f"output_head_dict.{k}.0.bias"] = (D,)
            out[f"output_head_dict.{k}.1.weight"] = (D, D)
            out[f"output_head_dict.{k}.1.bias"] = (D,)
            out[f"output_head_dict.{k}.3.weight"] = (d, D)
            out[f"output_head_dict.{k}.3.bias"] = (d,)
    for stack, n in (("encoder", n_enc_layer), ("decoder", n_dec_layer)):
        for i in range(n):
            p = f"{stack}.layers.{i}."
            out[p + "self_attn.in_proj_weight"] = (3 * D, D)
            out[p + "self_attn.in_proj_bias"] = (3 * D,)
            out[p + "self_attn.out_proj.weight"] = (D, D)
            out[p + "self_attn.out_proj.bias"] = (D,)
            out[p + "linear1.weight"] = (F, D)
            out[p + "linear1.bias"] = (F,)
            out[p + "linear2.weight"] = (D, F)
            out[p + "linear2.bias"] = (D,)
            for nm in ("norm1", "norm2"):
                out[p + nm + ".weight"] = (D,)
                out[p + nm + ".bias"] = (D,)
        out[f"{stack}.norm.weight"] = (D,)
        out[f"{stack}.norm.bias"] = (D,)
    return out


def validate_mtm_state_dict(sd: Mapping[str, torch.Tensor], data_shapes, n_embd: int, n_enc_layer: int, n_dec_layer: int,
                            traj_length: int) -> None:
    """Raise ``CheckpointError`` naming every missing / unexpected key and every shape mismatch (all at once, so a checkpoint
    trained with another config is diagnosed in one message instead of failing inside the engine)."""
    want = expected_mtm_keys(data_shapes, n_embd, n_enc_layer, n_dec_layer, traj_length)
    missing = sorted(set(want) - set(sd))
    extra = sorted(set(sd) - set(want))
    bad = [f"{k}: {tuple(sd[k].shape)} != {want[k]}" for k in want if k in sd and tuple(sd[k].shape) != want[k]]
    if missing or extra or bad:
        parts = []
        if missing:
            parts.append(f"missing {missing[:6]}{' ...' if len(missing) > 6 else ''} ({len(missing)})")
        if extra:
            parts.append(f"unexpected {extra[:6]}{' ...' if len(extra) > 6 else ''} ({len(extra)})")
        if bad:
            parts.append(f"shape mismatch {bad[:6]}{' ...' if len(bad) > 6 else ''} ({len(bad)})")
        raise CheckpointError("MTM checkpoint does not match the model config: " + "; ".join(parts))
    for k, v in sd.items():
        if not torch.isfinite(v.float()).all():
            raise CheckpointError(f"MTM checkpoint: non-finite values in {k}")


def load_mtm_checkpoint(path: str, trust_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """``torch.load(path)["model"]`` (learner.py:33-35) as fp32 CPU tensors; a DistributedDataParallel ``module.`` prefix is dropped.
    Also accepts a bare state dict (what ``torch.save(omtm.state_dict())`` writes)."""
    blob = _torch_load(path, trust_pickle)
    if isinstance(blob, Mapping) and "model" in blob and isinstance(blob["model"], Mapping):
        sd = blob["model"]
    elif isinstance(blob, Mapping) and blob and all(isinstance(v, torch.Tensor) for v in blob.values()):
        sd = blob
    else:
        raise CheckpointError(f"{path}: expected a dict with a 'model' state dict (finetune.py:319-326), got {type(blob).__name__}"
                              + (f" with keys {sorted(blob)[:8]}" if isinstance(blob, Mapping) else ""))
    sd = _strip_prefix(sd, "module.")
    return {k: v.detach().to(torch.float32).contiguous() for k, v in sd.items()}


def checkpoint_step(path: str, trust_pickle: bool = False) -> Optional[int]:
    """The ``step`` the trainer stored next to the weights (finetune.py:323), or None."""
    blob = _torch_load(path, trust_pickle)
    s = blob.get("step") if isinstance(blob, Mapping) else None
    return int(s) if s is not None else None


def load_iql_checkpoint(path: str, trust_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """The TwinQ state dict (``q1.net.{0,2,4}.*``, ``q2.net.{0,2,4}.*``) out of ``iql_<step>.pt`` (model.py:310-320).  The value
    function, actor and optimiser states in the file belong to training and are not on the planning path."""
    blob = _torch_load(path, trust_pickle)
    if not isinstance(blob, Mapping) or "qf" not in blob:
        raise CheckpointError(f"{path}: expected ImplicitQLearning.state_dict() with a 'qf' entry (model.py:310-320)")
    sd = _strip_prefix(blob["qf"], "module.")
    need = [f"q{q}.net.{l}.{p}" for q in (1, 2) for l in (0, 2, 4) for p in ("weight", "bias")]
    missing = [k for k in need if k not in sd]
    if missing:
        raise CheckpointError(f"{path}: TwinQ state dict lacks {missing[:4]} ... ({len(missing)} keys); only the reference's default "
                              "two-hidden-layer critic is supported")
    return {k: v.detach().to(torch.float32).contiguous() for k, v in sd.items() if k in need}


class _StatsUnpickler(pickle.Unpickler):
    """The cache holds instances of ``research.omtm.datasets.base.DataStatistics``; resolve that name to this package's mirror so
    the file loads without the reference on the path.  Everything else is restricted to numpy array reconstruction."""

    _ALLOWED = {("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"), ("numpy", "dtype"),
                ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("collections", "OrderedDict")}

    def find_class(self, module, name):
        if name == "DataStatistics":
            return DataStatistics
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"trajectory-statistics cache refers to {module}.{name}, which is not part of the format")


def load_trajectory_statistics(path: str) -> Dict[str, DataStatistics]:
    """``/tmp/d4rl/d4rl_statistics_<name>.pkl`` -> {"states", "actions", "rewards", "returns"} -> DataStatistics, with the
    reference's ``values`` -> ``returns`` rename (sequence_dataset.py:372-378)."""
    with open(path, "rb") as f:
        raw = _StatsUnpickler(io.BytesIO(f.read())).load()
    if not isinstance(raw, Mapping):
        raise CheckpointError(f"{path}: expected a dict of DataStatistics")
    out = dict(raw)
    if "values" in out:
        out["returns"] = out.pop("values")
    for k in MODALITIES:
        if k not in out:
            raise CheckpointError(f"{path}: no statistics for {k!r} (has {sorted(out)})")
        if not isinstance(out[k], DataStatistics):
            raise CheckpointError(f"{path}: entry {k!r} is {type(out[k]).__name__}, not DataStatistics")
    return {k: out[k] for k in MODALITIES}


def tokenizer_manager_from_statistics(stats: Mapping[str, DataStatistics], normalize: bool = True) -> TokenizerManager:
    """What ``ContinuousTokenizer.create`` builds per modality from the dataset statistics (continuous.py:42-62): std < 0.1 is
    replaced by 1, actions are never normalised."""
    toks = {}
    for k in MODALITIES:
        s = stats[k]
        s = DataStatistics(np.array(s.mean), np.array(s.std), np.array(s.min), np.array(s.max))  # create() edits std in place: keep the caller's copy
        std = s.std
        std[std < 0.1] = 1
        toks[k] = ContinuousTokenizer(s.mean, std, s, normalize=(False if k == "actions" else normalize))
    return TokenizerManager(toks)


def data_shapes_from_statistics(stats: Mapping[str, DataStatistics]) -> Dict[str, Tuple[int, int]]:
    """(tokens per step, feature dim) per modality, in the key order the model expects."""
    return {k: (1, int(np.asarray(stats[k].mean).reshape(-1).shape[0])) for k in MODALITIES}


def build_learner(cfg, model_config, mtm_path: str, stats_path: str, *, iql_path: Optional[str] = None, obs_mean=None, obs_std=None,
                  zeroshot: bool = False, max_envs: int = 1, env=None, trust_pickle: bool = False):
    """The three files -> a ready planner: what finetune.py:176-224 / unseen.py:176-224 assemble from the dataset, the pretrained
    path and the IQL trainer, without the dataset or the trainer.  ``obs_mean`` / ``obs_std`` default to the states statistics
    (the replay buffer's normaliser, replay_buffer.py, is computed from the same observations)."""
    stats = load_trajectory_statistics(stats_path)
    shapes = data_shapes_from_statistics(stats)
    sd = load_mtm_checkpoint(mtm_path, trust_pickle)
    validate_mtm_state_dict(sd, shapes, model_config.n_embd, model_config.n_enc_layer, model_config.n_dec_layer, cfg.traj_length)
    tm = tokenizer_manager_from_statistics(stats)
    om = stats["states"].mean if obs_mean is None else obs_mean
    os_ = stats["states"].std if obs_std is None else obs_std
    discrete_map = {k: False for k in shapes}
    if zeroshot:
        from .zeroshot_learner import Learner as ZLearner
        L = ZLearner(cfg, env, shapes, model_config, None, om, os_, tm, discrete_map, max_envs=max_envs)
    else:
        from .learner import Learner as FLearner
        L = FLearner(cfg, env, shapes, model_config, None, om, os_, tm, discrete_map, max_envs=max_envs)
    L.mtm.load_state_dict(sd)
    if iql_path is not None and hasattr(L, "iql"):
        L.iql.qf.load_state_dict(load_iql_checkpoint(iql_path, trust_pickle))
    return L
