"""Batch sharding over the GPUs of one node.

Sequences are independent through all T steps, so the evaluation batch is split contiguously
over ranks (one process per GPU), with exactly two collectives per job: one broadcast of the
weights at start and one gather of the final tokens at the end.  There is no per-step
collective.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the plumbing.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process -> (0, 0, 1))."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init_from_env(backend: str | None = None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int):
    """Contiguous [lo, hi) slice of `total` sequences owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_model(model: torch.nn.Module, src: int = 0) -> int:
    """One-time weight broadcast: every state-dict tensor from `src` to all ranks.  Returns bytes."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    nbytes = 0
    for _, t in model.state_dict().items():
        buf = torch.view_as_real(t) if t.is_complex() else t
        dist.broadcast(buf, src=src)
        nbytes += buf.numel() * buf.element_size()
    return nbytes


def gather_tokens(tokens: torch.Tensor, dst: int = 0):
    """Gather the per-rank token matrices [B_r, L] on `dst` as int8 (23 symbols fit a byte).

    Returns the concatenated [sum B_r, L] tensor on `dst`, None elsewhere.  Ranks may hold
    different B_r (ragged shards): sizes are exchanged first.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tokens
    world, rank = dist.get_world_size(), dist.get_rank()
    small = tokens.to(torch.int8).contiguous()
    sizes = [torch.zeros(1, dtype=torch.int64, device=small.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([small.shape[0]], dtype=torch.int64, device=small.device))
    sizes = [int(s.item()) for s in sizes]
    bmax = max(sizes)
    pad = torch.zeros((bmax, small.shape[1]), dtype=torch.int8, device=small.device)
    pad[: small.shape[0]] = small
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[:n] for o, n in zip(out, sizes)], dim=0).to(tokens.dtype)
