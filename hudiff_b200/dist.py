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


def pack_state(model: torch.nn.Module):
    """(flat float32 blob, [(name, tensor-view-into-the-state-dict, offset, numel)]) of every state-dict tensor.

    Complex buffers (the RoPE tables) travel as interleaved (re, im) float pairs."""
    views, total = [], 0
    for name, t in model.state_dict().items():
        buf = torch.view_as_real(t) if t.is_complex() else t
        if not buf.dtype.is_floating_point:
            raise TypeError(f"state-dict tensor {name} has dtype {buf.dtype}; the denoisers hold floating-point tensors only")
        views.append((name, buf, total, buf.numel()))
        total += buf.numel()
    ref = views[0][1]
    blob = torch.empty(total, dtype=torch.float32, device=ref.device)
    for _, buf, off, n in views:
        blob[off:off + n] = buf.reshape(-1).to(torch.float32)
    return blob, views


def broadcast_model(model: torch.nn.Module, src: int = 0) -> int:
    """The ONE weight broadcast of a job: the whole state dict as a single packed fp32 blob (one NCCL call instead of
    ~470 per-tensor ones), unpacked in place on the receiving ranks.  Returns the bytes broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    blob, views = pack_state(model)
    dist.broadcast(blob, src=src)
    if dist.get_rank() != src:
        with torch.no_grad():
            for _, buf, off, n in views:
                buf.copy_(blob[off:off + n].view(buf.shape).to(buf.dtype))
    return blob.numel() * blob.element_size()


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
