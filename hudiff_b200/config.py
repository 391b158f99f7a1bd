"""Model hyper-parameters of the two HuDiff denoisers and the checkpoint-config shim.

The shipped training configs are the shape source of truth for the released
checkpoints (configs/antibody_train.yml:3-24, configs/heavy_train.yml:3-21 in
the reference); the released `.pt` files carry the same dict pickled as an
`easydict.EasyDict`.  `EasyDict` below is registered under the module name
`easydict` when that package is absent so `torch.load(..., weights_only=False)`
can unpickle reference checkpoints.
"""
from __future__ import annotations

import sys
import types


class EasyDict(dict):
    """dict with recursive attribute access (what the reference's configs are)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
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

    def __reduce__(self):
        return (EasyDict, (dict(self),))


def install_easydict_shim() -> None:
    """Make `import easydict` resolve to this shim if the real package is missing."""
    if "easydict" in sys.modules:
        return
    try:
        import easydict  # noqa: F401
    except ImportError:
        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m


# configs/antibody_train.yml:3-24
AB_RELEASE = dict(
    n_tokens=23, d_embedding=256, d_model=256, n_encoder_layers=6, aa_kernel_size=7, r=128,
    n_side=3, s_embedding=4, s_model=256, n_region=7, r_embedding=4, r_model=256,
    n_pos_model=256, max_len=291, sum_d_model=768, dual_layers=6, att_model=512,
    dim_feedforward=256, nhead=8, cs_layers=5, dropout=0.2, activation="gelu")

# configs/heavy_train.yml:3-21
NB_RELEASE = dict(
    n_tokens=23, d_embedding=256, d_model=256, n_encoder_layers=6, aa_kernel_size=7, r=128,
    n_region=7, r_embedding=4, r_model=256, n_pos_model=256, max_len=152, sum_d_model=512,
    dual_layers=6, att_model=512, dim_feedforward=256, nhead=8, cs_layers=5, dropout=0.5,
    activation="gelu")

# Reduced-depth variants used by fast parity tests (same widths -> same kernels,
# fewer layers so the CPU oracle finishes in seconds).
AB_SMALL = dict(AB_RELEASE, n_encoder_layers=2, dual_layers=2, cs_layers=1)
NB_SMALL = dict(NB_RELEASE, n_encoder_layers=2, dual_layers=2, cs_layers=1)


def release_config(kind: str, **over) -> EasyDict:
    """EasyDict in the layout the reference stores inside its checkpoints."""
    if kind == "ab":
        return EasyDict(name="trans_oadm", model=dict(AB_RELEASE, **over))
    if kind == "nb":
        return EasyDict(name="nano", model=dict(NB_RELEASE, **over))
    raise ValueError(kind)
