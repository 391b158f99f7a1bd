"""Import the UNMODIFIED reference model classes in the build container.  TEST INFRASTRUCTURE.

/root/reference depends on packages that are not installed (sequence_models,
abnumber, anarci, pymol, easydict, Bio, lmdb ...).  This module pre-registers
minimal stand-ins in `sys.modules` so that `model.encoder.model.AntiTFNet` and
`model.nanoencoder.model.NanoAntiTFNet` import and run on CPU unchanged.  It is
used only by `oracle/make_golden.py` (to generate the committed golden vectors)
and by CPU tests that are skipped when /root/reference is absent (the GPU box).

`sequence_models` is restated from its published algorithm (see the header of
`oracle/hudiff_oracle.py`); everything else is an inert placeholder.
"""
from __future__ import annotations

import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get("HUDIFF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "encoder"))


def _mod(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _PFF(nn.Module):
    def __init__(self, d_in, d_out, rank=None):
        super().__init__()
        assert rank is None
        self.conv = nn.Conv1d(d_in, d_out, 1)

    def forward(self, x):
        return self.conv(x.transpose(1, 2)).transpose(1, 2)


class _MaskedConv1d(nn.Conv1d):
    def __init__(self, cin, cout, kernel_size, stride=1, dilation=1, groups=1, bias=True):
        super().__init__(cin, cout, kernel_size, stride=stride, dilation=dilation, groups=groups,
                         bias=bias, padding=dilation * (kernel_size - 1) // 2)

    def forward(self, x, input_mask=None):
        if input_mask is not None:
            x = x * input_mask
        return super().forward(x.transpose(1, 2)).transpose(1, 2)


class _ByteNetBlock(nn.Module):
    def __init__(self, d_in, d_h, d_out, kernel_size, dilation=1, groups=1, causal=False,
                 activation="relu", rank=None):
        super().__init__()
        assert not causal
        act = nn.ReLU if activation == "relu" else nn.GELU
        self.conv = _MaskedConv1d(d_h, d_h, kernel_size=kernel_size, dilation=dilation, groups=groups)
        self.sequence1 = nn.Sequential(nn.LayerNorm(d_in), act(), _PFF(d_in, d_h),
                                       nn.LayerNorm(d_h), act())
        self.sequence2 = nn.Sequential(nn.LayerNorm(d_h), act(), _PFF(d_h, d_out))

    def forward(self, x, input_mask=None):
        return x + self.sequence2(self.conv(self.sequence1(x), input_mask=input_mask))


class EasyDict(dict):
    """Attribute-access dict (stand-in for easydict.EasyDict), recursive."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


_installed = False


def install():
    """Register the stubs and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    _mod("sequence_models")
    _mod("sequence_models.layers", PositionFeedForward=_PFF, DoubleEmbedding=nn.Module)
    _mod("sequence_models.convolutional", ByteNetBlock=_ByteNetBlock, MaskedConv1d=_MaskedConv1d)
    _mod("pymol", cmd=None)
    _mod("abnumber", Chain=object)
    _mod("easydict", EasyDict=EasyDict)
    _mod("model.nanoencoder.abnativ_scoring", get_abnativ_nativeness_scores=None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def load_models():
    """Returns (AntiTFNet, NanoAntiTFNet) classes of the reference."""
    install()
    from model.encoder.model import AntiTFNet          # noqa: E402
    from model.nanoencoder.model import NanoAntiTFNet  # noqa: E402
    return AntiTFNet, NanoAntiTFNet
