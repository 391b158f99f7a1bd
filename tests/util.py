"""Shared helpers for the test-suite (oracle side + seeded weights)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from hudiff_b200 import AB_RELEASE, NB_RELEASE, AntiTFNetB200, NanoAntiTFNetB200  # noqa: E402


def make_model(kind, cfg=None, seed=0):
    """Mirror module with default nn init under torch.manual_seed(seed) (== the reference's)."""
    cfg = cfg or (AB_RELEASE if kind == "ab" else NB_RELEASE)
    state = torch.get_rng_state()
    torch.manual_seed(seed)
    m = (AntiTFNetB200 if kind == "ab" else NanoAntiTFNetB200)(**cfg)
    torch.set_rng_state(state)
    return m, dict(cfg)


def weight_checksum(sd):
    s = a = 0.0
    for k in sorted(sd):
        v = sd[k]
        v = torch.view_as_real(v) if v.is_complex() else v
        s += float(v.double().sum())
        a += float(v.double().abs().sum())
    return np.array([s, a])


def rel_err(a, b):
    """max-norm relative error used for the logits tolerance."""
    return float((a - b).abs().max() / b.abs().max())
