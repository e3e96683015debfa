"""Multi-GPU plumbing for the planners: one process per GPU, ``torch.distributed`` over NCCL (gloo on CPU for tests).

The path shards in two ways:

  * environments / independent plans -- embarrassingly parallel, no collective at all (``bench.py`` default);
  * candidates of ONE plan -- rank g owns global candidates [lo, hi) (``shard_range``); pass 1 (B = 1) is replicated,
    noise is indexed by global candidate id, and the only exchange is an all-gather of one per-shard record of
    ``PARTIAL_FLOATS`` floats followed by a log-sum-exp merge.  Production transport: ``connect_exchange`` wires the engines'
    peer-mapped exchange buffers once (CUDA IPC handles travel through ``torch.distributed``), after which
    ``engine.plan(..., exchange=True)`` does selection + all-gather + merge in ONE kernel over NVLink peer memory, inside the
    plan's CUDA graph.  ``gather_partials`` (an NCCL all-gather launched from the host) + ``m3pc_merge_partials`` is the plain
    transport kept for comparison and for the gloo CPU tests.
  * weights: ``broadcast_parameters`` sends rank 0's parameters to every rank in one flat NCCL broadcast at load.

``merge_partials_host`` is the numpy statement of that merge; tests use it to check the device kernel and the gloo path.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch
import torch.distributed as dist

from ._native import PARTIAL_FLOATS


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Read RANK / LOCAL_RANK / WORLD_SIZE (torchrun) and join the process group if WORLD_SIZE > 1."""
    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n_total candidates: the first (n_total % world) ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_bytes(payload: bytes, group=None) -> list:
    """Every rank's ``payload`` in rank order (host side plumbing: IPC handles at connect time)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [payload]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, payload, group=group)
    return out


def connect_exchange(engine, group=None) -> Tuple[int, int]:
    """Wire ``engine`` (one per process, one process per GPU of ONE node) into the peer exchange of its process group: gather
    the CUDA IPC handles of all exchange buffers and open them (``m3pc_exchange_connect``).  Returns (rank, world)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    handle, _ = engine.exchange_local()
    engine.exchange_connect(rank, world, ipc_handles=all_gather_bytes(handle, group))
    if dist.is_initialized() and world > 1:
        dist.barrier(group=group)  # nobody plans before everybody has mapped everybody (the buffers were just zeroed)
    return rank, world


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """north_star: "weights are broadcast once".  All parameters and buffers of ``module`` leave rank ``src`` as ONE flat
    tensor (one NCCL broadcast over NVLink; gloo on CPU) and are copied back into place on the receivers.  Returns the number
    of bytes broadcast.  Call before the first plan (the engine packs its bf16 arena from these tensors on first use)."""
    tensors = [p.data for p in module.parameters()] + [b for b in module.buffers()]
    if not tensors:
        return 0
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(t.numel() * t.element_size() for t in tensors)
    dev = tensors[0].device
    flat = torch.cat([t.detach().reshape(-1).to(torch.float32) for t in tensors]).to(dev)
    dist.broadcast(flat, src=src, group=group)
    if dist.get_rank(group) != src:
        o = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                t.copy_(flat[o:o + n].reshape(t.shape).to(t.dtype))
                o += n
        if hasattr(module, "mark_dirty"):
            module.mark_dirty()
    return int(flat.numel() * 4)


def gather_partials(record: torch.Tensor) -> torch.Tensor:
    """All-gather one (PARTIAL_FLOATS,) record per rank -> (world, PARTIAL_FLOATS), same device as the input."""
    if record.numel() != PARTIAL_FLOATS:
        raise ValueError(f"record must have {PARTIAL_FLOATS} floats")
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return record.reshape(1, PARTIAL_FLOATS).clone()
    out = torch.empty(dist.get_world_size(), PARTIAL_FLOATS, dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(out, record.reshape(1, PARTIAL_FLOATS).contiguous())
    return out


def make_partial_host(J: np.ndarray, a0: np.ndarray, q: np.ndarray, temperature: float, cand_offset: int) -> np.ndarray:
    """numpy statement of the record ``select_kernel`` emits for one shard (include/m3pc.h, M3PC_PARTIAL_FLOATS)."""
    J = np.asarray(J, np.float32)
    a0 = np.asarray(a0, np.float32)
    A = a0.shape[1]
    rec = np.zeros(PARTIAL_FLOATS, np.float32)
    m = J.max()
    w = np.exp((J - m) * np.float32(temperature)).astype(np.float32)
    key = w / np.asarray(q, np.float32)
    ki, mi = int(np.argmax(key)), int(np.argmax(J))
    rec[0], rec[1], rec[2] = m, w.sum(dtype=np.float32), m
    rec[3:4] = np.array([cand_offset + mi], np.int32).view(np.float32)
    rec[4] = key[ki]
    rec[5:6] = np.array([cand_offset + ki], np.int32).view(np.float32)
    rec[6] = len(J)
    rec[8:8 + A] = (w[:, None] * a0).sum(axis=0, dtype=np.float32)
    rec[8 + A:8 + 2 * A] = a0[ki]
    return rec


def merge_partials_host(records: np.ndarray, act_dim: int, temperature: float):
    """numpy statement of ``m3pc_merge_partials``: returns (eval_action, sample_action, argmax idx, sampled idx)."""
    R = np.asarray(records, np.float32).reshape(-1, PARTIAL_FLOATS)
    m = R[:, 0].max()
    sc = np.exp((R[:, 0] - m) * np.float32(temperature)).astype(np.float32)
    Z = (R[:, 1] * sc).sum(dtype=np.float32)
    U = (R[:, 8:8 + act_dim] * sc[:, None]).sum(axis=0, dtype=np.float32)
    kg = int(np.argmax(R[:, 4] * sc))
    jg = int(np.argmax(R[:, 2]))
    idx = R[:, [3, 5]].copy().view(np.int32)
    return U / Z, R[kg, 8 + act_dim:8 + 2 * act_dim].copy(), int(idx[jg, 0]), int(idx[kg, 1])
