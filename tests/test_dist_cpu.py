"""world_size-2 gloo test of the sharding plumbing (the N>1 path of bench.py), on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hudiff_b200 import dist as hd
from hudiff_b200.synthetic import synthetic


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, _, w = hd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # weight broadcast: rank 1 starts from different weights and must end up with rank 0's
    torch.manual_seed(rank)
    lin = torch.nn.Linear(8, 8)
    lin.register_buffer("rope", torch.polar(torch.ones(4, 2), torch.randn(4, 2)))
    calls = []
    real_broadcast = dist.broadcast
    dist.broadcast = lambda *a, **k: (calls.append(1), real_broadcast(*a, **k))[1]
    nbytes = hd.broadcast_model(lin, src=0)
    dist.broadcast = real_broadcast
    assert len(calls) == 1                                   # ONE packed blob, not one collective per tensor
    torch.manual_seed(0)
    ref = torch.nn.Linear(8, 8)
    ref_rope = torch.polar(torch.ones(4, 2), torch.randn(4, 2))
    assert torch.equal(lin.weight, ref.weight) and torch.equal(lin.bias, ref.bias)
    assert lin.rope.dtype == torch.complex64 and torch.equal(lin.rope, ref_rope)
    assert nbytes == (64 + 8) * 4 + 4 * 2 * 2 * 4
    # ragged batch shards (total = 5 over 2 ranks -> 3 + 2), gathered in order on rank 0
    tokens, _, _, _ = synthetic("nb", total, 4, seed=7)
    lo, hi = hd.shard_range(total, rank, world)
    got = hd.gather_tokens(tokens[lo:hi].clone(), dst=0)
    if rank == 0:
        assert got.dtype == tokens.dtype and torch.equal(got, tokens)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_broadcast_gather_world2():
    mp.spawn(_worker, args=(2, _free_port(), 5), nprocs=2, join=True)


def test_shard_range_partitions_everything():
    for total in (0, 1, 5, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [hd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
