#!/usr/bin/env python
"""Diagnostic: identical replicas in one batch must give identical activations.  Localises the first buffer where
sequences of a batch of identical inputs differ (forward, dropout off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hudiff_b200 import NB_RELEASE, AB_RELEASE
from hudiff_b200.synthetic import synthetic
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import make_model

for kind, B in (("nb", 64), ("nb", 1024), ("ab", 64)):
    base = NB_RELEASE if kind == "nb" else AB_RELEASE
    cfg = dict(base, n_encoder_layers=1, dual_layers=1, cs_layers=1, dropout=0.0)
    model, cfg = make_model(kind, cfg, seed=13)
    model = model.cuda().set_precision("fp16")
    tokens, region, chain, order = synthetic(kind, 1, 1, seed=2)
    tokens, region = tokens.expand(B, -1).contiguous(), region.expand(B, -1).contiguous()
    if chain is not None:
        chain = torch.cat([chain[:1].expand(B), chain[1:].expand(B)]).contiguous()
    y = model(tokens.cuda(), region.cuda(), None if chain is None else chain.cuda())
    L = y.shape[1]
    print(kind, B, "logits spread", float((y - y[0:1]).abs().max()))
    eng = model.engine("cuda")
    for name in ("E32", "opE", "hE1", "hE2", "opD", "hD1", "hD2", "X32", "qkv", "o16", "A32", "f16"):
        try:
            buf = eng.debug_buffer(B, name).float().view(B, L, -1)
        except Exception as e:
            print(name, "n/a", e); continue
        d = (buf - buf[0:1]).abs().amax((1, 2))
        bad = torch.nonzero(d > 0).flatten()
        print(f"  {name:5s} max spread {float(d.max()):.3e}  differing sequences {len(bad)}  first {bad[:6].tolist()}")
