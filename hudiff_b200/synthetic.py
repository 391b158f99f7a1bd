"""Synthetic benchmark / test inputs of the shapes BASELINE.json names.

HuDiff-Ab: B independent paired sequences on the 152+139 IMGT slot grid, HuDiff-Nb: B VHH
sequences on the 152-slot grid.  Residues are uniform over the 20 amino acids; pad ('-', 21)
follows a fixed template that leaves a typical number of real residues; the region ids are
the reference's fixed tables; `T` framework slots per sequence are masked (22) and their draw
order is the per-sequence sampling order (int32 [B, T]).
"""
from __future__ import annotations

import numpy as np
import torch

from . import tables as tb

PAD, MSK = 21, 22


def _heavy_pad_template(rng, n):
    """bool [n, 152]: True where the slot is a gap.  CDR3 insertion slots 111..134 keep only
    the outer k~U{0..8} on each side; IMGT 10 (slot 9) and IMGT 73 (slot 72) are gaps."""
    pad = np.zeros((n, tb.H_LEN), dtype=bool)
    pad[:, [9, 72]] = True
    k = rng.integers(0, 9, size=n)
    for i in range(n):
        pad[i, 111 + k[i]: 135 - k[i]] = True
    return pad


def _light_pad_template(n):
    pad = np.zeros((n, tb.L_LEN), dtype=bool)
    pad[:, 111:123] = True          # IMGT 111A..112A insertions
    pad[:, 29:35] = True            # short CDR-L1
    return pad


def _draw_order(rng, cand_mask, T):
    B = cand_mask.shape[0]
    order = np.zeros((B, T), dtype=np.int32)
    for i in range(B):
        cand = np.flatnonzero(cand_mask[i])
        if len(cand) < T:
            raise ValueError(f"T={T} exceeds the {len(cand)} maskable slots of sequence {i}")
        order[i] = rng.permutation(cand)[:T]
    return order


def synthetic_ab(B: int, T: int = 50, seed: int = 0):
    """-> tokens int64 [B,291], region int64 [B,291], chain int64 [2B], order int32 [B,T]."""
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, 20, size=(B, tb.AB_LEN))
    pad = np.concatenate([_heavy_pad_template(rng, B), _light_pad_template(B)], axis=1)
    tok[pad] = PAD
    frame = np.array(tb.HEAVY_CDR_KABAT_NO_VERNIER + tb.LIGHT_CDR_KABAT_NO_VERNIER) == 0
    order = _draw_order(rng, frame[None, :] & ~pad, T)
    np.put_along_axis(tok, order.astype(np.int64), MSK, axis=1)
    region = np.tile(np.array(tb.HEAVY_REGION_INDEX + tb.LIGHT_REGION_INDEX), (B, 1))
    chain = np.concatenate([np.zeros(B, dtype=np.int64), rng.integers(1, 3, size=B)])
    return (torch.from_numpy(tok.astype(np.int64)), torch.from_numpy(region.astype(np.int64)),
            torch.from_numpy(chain.astype(np.int64)), torch.from_numpy(order))


def synthetic_nb(B: int, T: int = 50, seed: int = 0, inpaint: bool = False):
    """-> tokens int64 [B,152], region int64 [B,152], None, order int32 [B,T]."""
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, 20, size=(B, tb.H_LEN))
    pad = _heavy_pad_template(rng, B)
    tok[pad] = PAD
    table = tb.INPAINT_HEAVY_CDR_INDEX if inpaint else tb.HEAVY_CDR_INDEX
    frame = np.array(table) == 0
    order = _draw_order(rng, frame[None, :] & ~pad, T)
    np.put_along_axis(tok, order.astype(np.int64), MSK, axis=1)
    region = np.tile(np.array(tb.HEAVY_REGION_INDEX), (B, 1))
    return (torch.from_numpy(tok.astype(np.int64)), torch.from_numpy(region.astype(np.int64)),
            None, torch.from_numpy(order))


def synthetic(kind: str, B: int, T: int = 50, seed: int = 0):
    return synthetic_ab(B, T, seed) if kind == "ab" else synthetic_nb(B, T, seed)
