"""Host replay of the reference's CPU random streams (the `torch_cpu_replay` mode).

The reference forward is stochastic in eval mode (functional F.dropout) and the
loop samples with torch.multinomial.  On CPU both reduce to calls on torch's default generator:
    F.dropout(x, p)          == x * empty_like(x).bernoulli_(1 - p).div_(1 - p)
    torch.multinomial(pr, 1) == argmax(pr / empty_like(pr).exponential_(1))
The stream does not depend on the data, so all T steps can be drawn up front, in the
reference's call order, and shipped to the GPU as keep-bit planes and Exp(1) draws.  Only torch
can reproduce its own (MKL-backed) bernoulli/exponential kernels, so this module calls torch;
it performs no model arithmetic.
"""
from __future__ import annotations

import numpy as np
import torch

from .tables import H_LEN, L_LEN

N_SAMPLED = 22


def dropout_call_shapes(kind: str, cfg: dict, B: int):
    """[(shape, p)] of every functional dropout call of ONE forward, in the reference's order.

    Ab: model/encoder/model.py:173-178 (aa_encoder: per layer H then L, p = cfg.dropout) and
    :292-304 (DualConv: all heavy layers, then all light layers, p = 0.5).
    Nb: model/nanoencoder/model.py:166-169, 262-265.
    """
    p = float(cfg.get("dropout", 0.0))
    if not p > 0.0:
        return []
    d, D = cfg["d_model"], cfg["sum_d_model"]
    calls = []
    if kind == "ab":
        for _ in range(cfg["n_encoder_layers"]):
            calls += [((B, H_LEN, d), p), ((B, L_LEN, d), p)]
        calls += [((B, H_LEN, D), 0.5)] * cfg["dual_layers"]
        calls += [((B, L_LEN, D), 0.5)] * cfg["dual_layers"]
    else:
        calls += [((B, H_LEN, d), p)] * cfg["n_encoder_layers"]
        calls += [((B, H_LEN, D), 0.5)] * cfg["dual_layers"]
    return calls


def pack_keep_bits(mask: torch.Tensor) -> np.ndarray:
    """float/bool keep-mask [..., C] (C % 32 == 0) -> uint32 words, bit j of a word = column 32w + j."""
    m = (mask != 0).numpy().reshape(-1)
    return np.packbits(m, bitorder="little").view("<u4")


def draw_forward_bits(kind: str, cfg: dict, B: int) -> np.ndarray:
    """Consume torch's CPU generator exactly like one reference forward; return the packed bits."""
    words = [pack_keep_bits(torch.empty(shape, dtype=torch.float32).bernoulli_(1.0 - p))
             for shape, p in dropout_call_shapes(kind, cfg, B)]
    return np.concatenate(words) if words else np.zeros(0, dtype="<u4")


def draw_loop_streams(kind: str, cfg: dict, B: int, T: int, dropout: bool = True):
    """Streams of a T-step loop: (bits uint32 [T, words] or None, exp float32 [T, B, 22])."""
    bits, exps = [], []
    live = dropout and float(cfg.get("dropout", 0.0)) > 0.0
    for _ in range(T):
        if live:
            bits.append(draw_forward_bits(kind, cfg, B))
        exps.append(torch.empty((B, N_SAMPLED), dtype=torch.float32).exponential_(1.0))
    bits_t = torch.from_numpy(np.stack(bits).astype(np.uint32).view(np.int32)) if live else None
    return bits_t, torch.stack(exps)
