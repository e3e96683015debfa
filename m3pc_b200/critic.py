"""TwinQ critic -- parameter-compatible mirror of research/finetune_omtm/model.py:72-104 (MLP) and :146-171 (TwinQ).

Inside the fused planner the critic runs as part of ``m3pc_plan`` (K7).  Called on its own (``TwinQ.forward(state,
action)``, as ``critic_lambda_guiding`` does at learner.py:250-252) it runs the same fp32 CUDA kernels through the
C-ABI (``m3pc_gemm_fp32``); there is no CPU path.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from . import _native as nat


class Squeeze(nn.Module):
    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        return x.squeeze(dim=self.dim)


class MLP(nn.Module):
    """ReLU MLP whose ``state_dict`` keys are ``net.{0,2,4}.{weight,bias}`` like the reference's."""

    def __init__(self, dims, squeeze_output: bool = False):
        super().__init__()
        if len(dims) < 2:
            raise ValueError("MLP requires at least two dims (input and output)")
        layers = []
        for i in range(len(dims) - 2):
            layers += [nn.Linear(dims[i], dims[i + 1]), nn.ReLU()]
        layers.append(nn.Linear(dims[-2], dims[-1]))
        if squeeze_output:
            if dims[-1] != 1:
                raise ValueError("Last dim must be 1 when squeezing")
            layers.append(Squeeze(-1))
        self.net = nn.Sequential(*layers)

    def linears(self):
        return [m for m in self.net if isinstance(m, nn.Linear)]

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise NotImplementedError("TwinQ runs on the B200 engine only (no CPU path)")
        L = nat.lib()
        st = torch.cuda.current_stream().cuda_stream
        h = x.to(torch.float32).contiguous()
        lins = self.linears()
        for i, lin in enumerate(lins):
            out = torch.empty(h.shape[0], lin.out_features, device=h.device)
            flags = 4 if i + 1 < len(lins) else 0  # ReLU on hidden layers
            w, b = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
            nat.check(L.m3pc_gemm_fp32(h.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), h.shape[0], lin.out_features,
                                       lin.in_features, flags, st), "m3pc_gemm_fp32")
            h = out
        return h.squeeze(-1) if isinstance(self.net[-1], Squeeze) else h


class TwinQ(nn.Module):
    def __init__(self, state_dim: int, action_dim: int, obs_mean: torch.Tensor, obs_std: torch.Tensor, hidden_dim: int = 256, n_hidden: int = 2):
        super().__init__()
        if n_hidden != 2:
            raise NotImplementedError("the fused critic kernel sequence assumes two hidden layers (the reference default)")
        dims = [state_dim + action_dim, *([hidden_dim] * n_hidden), 1]
        self.q1 = MLP(dims, squeeze_output=True)
        self.q2 = MLP(dims, squeeze_output=True)
        self.obs_mean = obs_mean
        self.obs_std = obs_std
        self.hidden_dim = hidden_dim

    def both(self, state: torch.Tensor, action: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        state = (state - self.obs_mean.to(state.device)) / self.obs_std.to(state.device)
        sa = torch.cat([state, action], 1)
        return self.q1(sa), self.q2(sa)

    def forward(self, state: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        return torch.min(*self.both(state, action))
