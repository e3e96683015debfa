"""Per-modality tokenizers -- host-side mirror of the reference's API.

  DataStatistics      research/omtm/datasets/base.py:31-48
  ContinuousTokenizer research/omtm/tokenizers/continuous.py:31-94
  TokenizerManager    research/omtm/tokenizers/base.py:64-99

Only the continuous tokenizer exists: it is the only one any shipped config uses.  The planners
do not call these per step -- the CUDA engine normalises inside its embedding kernel and de-normalises inside its
scoring kernels -- but ``omtm.forward`` users (and checkpoints / configs written against the reference) get the
same objects with the same semantics.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import numpy as np
import torch


@dataclass
class DataStatistics:
    mean: np.ndarray
    std: np.ndarray
    min: np.ndarray
    max: np.ndarray

    def __post_init__(self):
        for f in ("mean", "std", "min", "max"):
            setattr(self, f, np.array(getattr(self, f), dtype=np.float32))
        assert self.mean.shape == self.std.shape == self.min.shape == self.max.shape
        assert np.all(self.min <= self.max)


class Tokenizer(torch.nn.Module):
    @property
    def discrete(self) -> bool:
        raise NotImplementedError


class ContinuousTokenizer(Tokenizer):
    def __init__(self, data_mean, data_std, stats: DataStatistics, normalize: bool = True):
        super().__init__()
        self._data_mean = torch.nn.Parameter(torch.tensor(np.asarray(data_mean), dtype=torch.float32), requires_grad=False)
        self._data_std = torch.nn.Parameter(torch.tensor(np.asarray(data_std), dtype=torch.float32), requires_grad=False)
        self.stats = stats
        self.normalize = normalize

    @classmethod
    def create(cls, key: str, train_dataset, normalize: bool = True) -> "ContinuousTokenizer":
        stats = train_dataset.trajectory_statistics()[key]
        data_mean, data_std = stats.mean, stats.std
        data_std[data_std < 0.1] = 1  # continuous.py:58 -- near-constant features are left un-scaled
        if key == "actions":
            return cls(data_mean, data_std, stats, normalize=False)
        return cls(data_mean, data_std, stats, normalize=normalize)

    @property
    def discrete(self) -> bool:
        return False

    def encode(self, trajectory: torch.Tensor) -> torch.Tensor:
        assert trajectory.dim() == 3
        if self.normalize:
            trajectory = (trajectory - self._data_mean.to(trajectory.device)) / self._data_std.to(trajectory.device)
        return trajectory.unsqueeze(2).to(torch.float32)

    def decode(self, trajectory):
        is_dist = hasattr(trajectory, "loc") and hasattr(trajectory, "sample")
        assert is_dist or trajectory.dim() == 4
        assert is_dist or trajectory.size(2) == 1
        if self.normalize:
            return trajectory.squeeze(2) * self._data_std.to(trajectory.device) + self._data_mean.to(trajectory.device)
        return trajectory


class TokenizerManager(torch.nn.Module):
    def __init__(self, tokenizers: Dict[str, Tokenizer]):
        super().__init__()
        self.tokenizers = torch.nn.ModuleDict(tokenizers)

    def encode(self, trajectories: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = {}
        for key, value in trajectories.items():
            if key in self.tokenizers.keys():
                out[key] = self.tokenizers[key].encode(value)
                assert len(out[key].shape) == 4
        return out

    def decode(self, tokenized: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {key: self.tokenizers[key].decode(value) for key, value in tokenized.items()}

    def engine_stats(self) -> Dict[str, Dict[str, np.ndarray]]:
        """mean/std per normalised modality, in the form ``PlanEngine.load_tokenizer_stats`` takes."""
        return engine_stats(self)


def engine_stats(manager) -> Dict[str, Dict[str, np.ndarray]]:
    """``TokenizerManager.engine_stats`` for this package's manager OR the reference's own
    (research/omtm/tokenizers/base.py:64-99 with ContinuousTokenizer, continuous.py:31-62): both keep ``_data_mean`` /
    ``_data_std`` / ``normalize`` per tokenizer.  Any other tokenizer type is rejected (only ContinuousTokenizer is configured
    by the shipped configs and supported by the engine)."""
    out = {}
    for k, tok in manager.tokenizers.items():
        if type(tok).__name__ != "ContinuousTokenizer" or not hasattr(tok, "_data_mean"):
            raise NotImplementedError(f"tokenizer for {k!r} is {type(tok).__name__}: only ContinuousTokenizer is supported")
        d = tok._data_mean.numel()
        if tok.normalize:
            out[k] = {"mean": tok._data_mean.detach().cpu().numpy(), "std": tok._data_std.detach().cpu().numpy()}
        else:
            out[k] = {"mean": np.zeros(d, np.float32), "std": np.ones(d, np.float32)}
    return out


def manager_from_stats(stats: Dict[str, Dict[str, np.ndarray]]) -> TokenizerManager:
    """Build the four ContinuousTokenizers from a mean/std/min/max dict (``m3pc_b200.synthetic.make_tokenizer_stats``)."""
    toks = {}
    for k in ("states", "actions", "rewards", "returns"):
        s = stats[k]
        toks[k] = ContinuousTokenizer(s["mean"], s["std"], DataStatistics(s["mean"], s["std"], s["min"], s["max"]), normalize=(k != "actions"))
    return TokenizerManager(toks)
